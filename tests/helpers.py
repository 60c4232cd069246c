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
