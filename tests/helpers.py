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
# ReLU-sign-aware gradient comparison
# ------------------------------------------------------------------------------------------------
def _flat(grads, names):
    return torch.cat([grads[n].detach().double().cpu().reshape(-1) for n in names])


def oracle_grads_with_forced(cfg, model, batch, forced):
    x = {k: v.double() for k, v in batch.x_dict.items()}
    model.zero_grad()
    with O.ReluTap(forced=forced):
        out = model(x, batch.edge_index_dict)
        loss = oracle_loss(cfg, out, batch.y.double(), batch.batch_size)
        loss.backward()
    return {n: (p.grad.clone() if p.grad is not None else torch.zeros_like(p)) for n, p in model.named_parameters()}


def relu_aware_gradient_check(cfg, model, batch, native_grads, fwd_rel_err, tol=TOL_FP32, max_candidates=400):
    """Returns (ok, info).  The native gradient must equal, per tensor within ``tol`` (norm-wise), the oracle's
    fp64 gradient for SOME ReLU sign pattern that differs from the oracle's own only at numerically ambiguous
    pre-activations: |x| < 16 * fwd_rel_err * rms(x), i.e. inside the native path's measured forward error.

    Why: ReLU makes d(loss)/d(params) discontinuous in the forward numerics.  An implementation whose forward
    pass is accurate to 1e-6 may legitimately put a pre-activation of +-1e-7 on the other side of zero, and that
    moves whole gradient tensors by 1e-4..1e-2 (plain PyTorch fp32 does the same against fp64).  Single flips act
    (to first order) additively, so the set of flipped elements is found by projecting the residual on each
    candidate's gradient delta and then VERIFIED with one exact oracle backward under the chosen pattern."""
    names = [n for n, _ in model.named_parameters()]
    x = {k: v.double() for k, v in batch.x_dict.items()}
    model.zero_grad()
    with O.ReluTap(record=True) as tap:
        out = model(x, batch.edge_index_dict)
        loss = oracle_loss(cfg, out, batch.y.double(), batch.batch_size)
        loss.backward()
    g0 = {n: (p.grad.clone() if p.grad is not None else torch.zeros_like(p)) for n, p in model.named_parameters()}

    def worst(ga):
        w = 0.0
        for n in names:
            nb = ga[n].norm().item()
            if nb == 0.0:
                if native_grads[n].abs().max().item() != 0.0:
                    return float("inf")
                continue
            w = max(w, rel_err(native_grads[n], ga[n]))
        return w

    e0 = worst(g0)
    if e0 <= tol:
        return True, {"flips": 0, "err": e0}
    # candidates: ambiguous pre-activations that carry gradient
    cands = []
    for i, pre in enumerate(tap.pre):
        if pre.grad is None:
            continue
        v = pre.detach()
        thr = 16.0 * max(fwd_rel_err, 1e-7) * v.pow(2).mean().sqrt().item()
        # the upstream gradient of relu(x) is what matters: grad wrt x is zero where the mask is off, so use |x| only
        idx = torch.nonzero(v.abs().reshape(-1) < thr).reshape(-1)
        for j in idx.tolist():
            cands.append((i, j, bool(v.reshape(-1)[j] > 0)))
    if not cands or len(cands) > max_candidates:
        return False, {"flips": None, "err": e0, "candidates": len(cands)}
    def merge(chosen):
        forced = {}
        for (i, j, cur) in chosen:
            idx, val = forced.get(i, (torch.empty(0, dtype=torch.long), torch.empty(0, dtype=torch.bool)))
            forced[i] = (torch.cat((idx, torch.tensor([j]))), torch.cat((val, torch.tensor([not cur]))))
        return forced

    # Greedy rounds: project the residual on every remaining candidate's gradient delta (taken around the pattern chosen
    # so far, so interactions between flips on one path are seen in the next round), keep those that explain it,
    # and VERIFY with an exact oracle backward under the chosen pattern.
    chosen, g_cur, e_cur = [], g0, e0
    for _ in range(4):
        r = _flat(native_grads, names) - _flat(g_cur, names)
        base = _flat(g_cur, names)
        picked = []
        for c in cands:
            if c in chosen:
                continue
            gj = oracle_grads_with_forced(cfg, model, batch, merge(chosen + [c]))
            d = _flat(gj, names) - base
            dn = d.dot(d).item()
            if dn > 0 and r.dot(d).item() / dn > 0.5:
                picked.append(c)
        if not picked:
            break
        chosen += picked
        g_cur = oracle_grads_with_forced(cfg, model, batch, merge(chosen))
        e_cur = worst(g_cur)
        if e_cur <= tol:
            break
    return e_cur <= tol, {"flips": len(chosen), "err": e_cur, "err_default_pattern": e0, "candidates": len(cands)}
