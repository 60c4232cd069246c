"""Shared helpers of the parity tests: the oracle is the checker, the native path the thing checked."""
import torch

import mshgnn_oracle as O
from ms_hgnn.synthetic import CONFIGS, build_model, make_batch

TOL_FP32 = 1e-4      # north_star: within 1e-4 relative (norm-wise, per tensor) in fp32 against the fp64 oracle


def rel_err(a: torch.Tensor, b: torch.Tensor) -> float:
    a = a.detach().double().cpu().flatten(); b = b.detach().double().cpu().flatten()
    nb = b.norm().item()
    if nb == 0.0:
        return (a - b).norm().item()
    return ((a - b).norm() / nb).item()


def oracle_model(cfg, hidden=128, layers=8, seed=0):
    """fp64 oracle whose weights are exactly representable in fp32 (so both sides see the same numbers)."""
    prev = torch.get_default_dtype()
    torch.set_default_dtype(torch.float64)
    try:
        m = build_model(cfg, hidden, layers, seed, module=O)
        with torch.no_grad():
            for p in m.parameters():
                p.copy_(p.float().double())
    finally:
        torch.set_default_dtype(prev)
    return m


def oracle_loss(cfg, out, y, B):
    if cfg.loss == "ce":
        return O.contact_ce_loss_exact(out.reshape(B, 8), y.reshape(B, 4))
    return O.mse_loss(out.reshape(B, -1), y.reshape(B, -1))


def oracle_run(cfg, model, batch):
    """fp64 forward + loss + backward of the oracle on a host batch; returns out, loss, {name: grad}."""
    x = {k: v.double() for k, v in batch.x_dict.items()}
    model.zero_grad()
    out = model(x, batch.edge_index_dict)
    loss = oracle_loss(cfg, out, batch.y.double(), batch.batch_size)
    loss.backward()
    grads = {n: (p.grad.clone() if p.grad is not None else torch.zeros_like(p)) for n, p in model.named_parameters()}
    return out.detach(), loss.detach(), grads


def oracle_fp32_floor(cfg, model, batch, ref_out, ref_grads):
    """Error of the SAME oracle op sequence run in plain PyTorch float32 against its float64 self.

    ReLU makes the gradient discontinuous in the forward numerics: an fp32 rounding that flips the sign
    of one pre-activation changes a whole gradient row.  No fp32 implementation (including the
    reference's own ops in fp32) can be closer to the fp64 oracle than this floor, so the parity bound
    for gradients is max(1e-4, 2 x floor) per tensor."""
    import copy
    m32 = copy.deepcopy(model).float()
    x = {k: v.float() for k, v in batch.x_dict.items()}
    m32.zero_grad()
    out = m32(x, batch.edge_index_dict)
    loss = oracle_loss(cfg, out, batch.y.float(), batch.batch_size)
    loss.backward()
    floor = {n: (rel_err(p.grad, ref_grads[n]) if p.grad is not None else 0.0) for n, p in m32.named_parameters()}
    return rel_err(out, ref_out), floor


# ------------------------------------------------------------------------------------------------
# oracle backward under a forced ReLU sign pattern
# ------------------------------------------------------------------------------------------------
def oracle_grads_with_forced(cfg, model, batch, forced):
    x = {k: v.double() for k, v in batch.x_dict.items()}
    model.zero_grad()
    with O.ReluTap(forced=forced):
        out = model(x, batch.edge_index_dict)
        loss = oracle_loss(cfg, out, batch.y.double(), batch.batch_size)
        loss.backward()
    return {n: (p.grad.clone() if p.grad is not None else torch.zeros_like(p)) for n, p in model.named_parameters()}


# ------------------------------------------------------------------------------------------------
# Gradient comparison under the ReLU sign pattern the native forward actually took
# ------------------------------------------------------------------------------------------------
def native_relu_pattern(model, B):
    """{oracle tap tag: BoolTensor [B * nodes_of_type, 128]} read back from the native training workspace
    (mshgnn_relu_mask_offset), plus the plan's liveness table.  Rows are re-ordered to the oracle's graph-major layout."""
    from ms_hgnn import _native as N
    eng = next(iter(model._engines.values()))
    spec, plan = eng.spec, eng.plan
    types, npg = spec["node_types"], spec["nodes_per_graph"]
    base = [sum(npg[:i]) for i in range(len(types))]
    S = sum(npg)
    need = plan.describe()["need"]                       # need[l][slot], l = 0..L
    torch.cuda.synchronize()
    ws = eng._ws
    shifts = torch.arange(32, device=ws.device, dtype=torch.int64)

    def read(layer):
        off, ns, bp = plan.relu_mask_offset(B, eng.mode, layer)
        words = ws[off:off + ns * bp * 16].view(torch.int32).view(ns, bp, 4)[:, :B].to(torch.int64) & 0xFFFFFFFF
        return ((words.unsqueeze(-1) >> shifts) & 1).bool().reshape(ns, B, 128).cpu()          # [slot][graph][channel]

    def rows_of(bits, first_slot, n):                    # -> [B * n, 128], row = graph * n + local node
        return bits[first_slot:first_slot + n].permute(1, 0, 2).reshape(B * n, 128)

    def live_of(l, first_slot, n):                       # [B * n] bool
        return torch.tensor([bool(need[l][first_slot + j]) for j in range(n)]).repeat(B)

    out = {}
    enc = read(-1)
    for t, name in enumerate(types):
        out[("enc", name)] = (rows_of(enc, base[t], npg[t]), live_of(0, base[t], npg[t]))
    mlp_t = spec["mlp_type"]
    for l in range(spec["num_layers"]):
        bits = read(l)
        for t, name in enumerate(types):
            if spec["morph_sym"] and t == mlp_t:
                out[("mlp", l, name)] = (rows_of(bits, S, npg[t]), live_of(l + 1, base[t], npg[t]))
            else:
                out[("conv", l, name)] = (rows_of(bits, base[t], npg[t]), live_of(l + 1, base[t], npg[t]))
    return out


def gradient_check_under_native_pattern(cfg, model, batch, native_model, native_grads, fwd_rel_err, tol=TOL_FP32, plain_grads=None):
    """Returns (ok, info).  ReLU makes d(loss)/d(params) discontinuous in the forward numerics: an implementation whose
    forward is accurate to 1e-6 may put a pre-activation of +-1e-7 on the other side of zero, which moves whole gradient
    tensors by 1e-4..1e-2 (plain PyTorch fp32 does the same against fp64).  So the native gradient is compared, strictly
    (``tol`` per tensor, norm-wise), with the oracle's fp64 gradient under the sign pattern the native forward took, after
    checking that this pattern differs from the oracle's own ONLY at numerically ambiguous pre-activations:
    |x| < 16 * max(forward error, 1e-7) * rms(x)."""
    names = [n for n, _ in model.named_parameters()]
    B = batch.batch_size
    pattern = native_relu_pattern(native_model, B)
    x = {k: v.double() for k, v in batch.x_dict.items()}
    with torch.no_grad(), O.ReluTap(record=True) as tap:
        model(x, batch.edge_index_dict)
    forced, flips, worst_ratio = {}, 0, 0.0
    for i, (pre, tag) in enumerate(zip(tap.pre, tap.tags)):
        if tag not in pattern:
            return False, {"error": f"no native pattern for oracle ReLU {tag}"}
        bits, live = pattern[tag]
        v = pre.detach()
        if tuple(bits.shape) != tuple(v.shape):
            return False, {"error": f"shape mismatch at {tag}: {tuple(bits.shape)} vs {tuple(v.shape)}"}
        diff = ((v > 0) != bits) & live[:, None]
        if diff.any():
            thr = 16.0 * max(fwd_rel_err, 1e-7) * v.pow(2).mean().sqrt().item()
            ratio = (v.abs()[diff].max().item() / thr) if thr > 0 else float("inf")
            worst_ratio = max(worst_ratio, ratio)
            idx = torch.nonzero(diff.reshape(-1)).reshape(-1)
            forced[i] = (idx, bits.reshape(-1)[idx])
            flips += int(idx.numel())
    if worst_ratio > 1.0:
        return False, {"error": "native ReLU pattern differs from the oracle at a NON-ambiguous pre-activation",
                       "worst |x| / threshold": worst_ratio, "flips": flips}
    g = oracle_grads_with_forced(cfg, model, batch, forced)
    worst, worst_name = 0.0, None
    for n in names:
        if g[n].norm().item() == 0.0:
            if native_grads[n].abs().max().item() != 0.0:
                return False, {"error": f"{n}: oracle gradient is exactly zero, native is not"}
            continue
        e = rel_err(native_grads[n], g[n])
        if e > worst:
            worst, worst_name = e, n
    info = {"flips": flips, "err": worst, "tensor": worst_name, "worst |x| / threshold": worst_ratio}
    if plain_grads is not None:
        # the PLAIN error too (oracle's own ReLU pattern, nothing forced): reported next to the forced one so that the
        # size of the sign-flip effect is visible; it equals the forced error when no pre-activation flipped
        pw, pn = 0.0, None
        for n in names:
            if plain_grads[n].norm().item() == 0.0:
                continue
            e = rel_err(native_grads[n], plain_grads[n])
            if e > pw:
                pw, pn = e, n
        info["plain_err"], info["plain_tensor"] = pw, pn
    return worst <= tol, info
