"""GPU tests of the reference-facing layer: Lightning-shaped modules, checkpoints, fused trainer."""
import os

import pytest
import torch

import mshgnn_oracle as O
from helpers import TOL_FP32, oracle_loss, oracle_model, rel_err
from ms_hgnn import morphology as M
from ms_hgnn.lightning_py.gnnLightning import (Heterogeneous_GNN_Lightning, HGNN_C2_Lightning_Reg, HGNN_K4_Lightning)
from ms_hgnn.lightning_py.gnnLightning_com import COM_HGNN_SYM_Lightning
from ms_hgnn.synthetic import CONFIGS, HeteroBatch, make_batch
from ms_hgnn.train import FusedTrainer

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
DEV = "cuda:0"


def test_reference_checkpoint_weights_through_the_native_path():
    """Real trained MI-HGNN weights + the real 20-graph batch (fixture from the reference's checkpoint):
    predictions within 1e-4 of the fp64 oracle, MSE/RMSE/L1 equal to the reference's pins to 4 decimals."""
    g = torch.load(os.path.join(GOLD, "mi_grf_ckpt.pt"), weights_only=False)
    B = g["x"]["base"].shape[0]
    batch = HeteroBatch(g["x"], g["edge_index"], g["y"], B)          # float64 tensors, as the reference feeds them
    mod = Heterogeneous_GNN_Lightning(g["hidden"], g["layers"], g["metadata"], batch, "adam", 1e-4, regression=True)
    missing = mod.load_state_dict({"model." + k: v for k, v in g["state_dict_f32"].items()})
    assert not missing.missing_keys and not missing.unexpected_keys
    mod = mod.to(DEV)
    b = batch.to(DEV)
    with torch.no_grad():
        y, y_pred = mod.step_helper_function(b)
        mod.calculate_losses_step(y, y_pred)
    assert y_pred.dtype == torch.float64 and tuple(y_pred.shape) == (B, 4)
    assert rel_err(y_pred.reshape(-1), g["oracle_out"].reshape(-1)) <= TOL_FP32
    pins = g["reference_pins"]
    assert abs(mod.mse_loss.item() - pins[0]) < 1e-4 * pins[0]
    assert abs(mod.rmse_loss.item() - pins[1]) < 1e-4 * pins[1]
    assert abs(mod.l1_loss.item() - pins[2]) < 1e-4 * pins[2]


def _oracle_adam_steps(cfg, om, batches, lr, steps):
    opt = torch.optim.Adam(om.parameters(), lr=lr)
    losses = []
    for i in range(steps):
        b = batches[i % len(batches)]
        opt.zero_grad()
        out = om({k: v.double() for k, v in b.x_dict.items()}, b.edge_index_dict)
        loss = oracle_loss(cfg, out, b.y.double(), b.batch_size)
        loss.backward()
        opt.step()
        losses.append(loss.item())
    return losses


@pytest.mark.parametrize("name", ["mini_cheetah-k4-contact", "a1-c2-grf", "solo12-k4-com"])
def test_training_steps_match_oracle_adam(name):
    """3 optimisation steps: (a) Lightning-shaped training_step + torch.optim.Adam through autograd,
    (b) the fused native trainer (forward -> loss -> backward -> fused Adam); both vs the fp64 oracle + Adam."""
    cfg = CONFIGS[name]
    tpl = M.TEMPLATES[cfg.template]
    B, L, lr, steps = 40, 4, 1e-3, 3
    batches = [make_batch(cfg, B, seed=s) for s in (1, 2)]
    om = oracle_model(cfg, layers=L, seed=7)
    sd = {"model." + k: v.float() for k, v in om.state_dict().items()}
    ref_losses = _oracle_adam_steps(cfg, om, batches, lr, steps)
    ref_params = {k: v.detach() for k, v in om.state_dict().items()}

    def make():
        kw = dict(symmetry_mode="MorphSym", group_operator_path=M.cfg_path(cfg.group))
        if name == "mini_cheetah-k4-contact":
            m = HGNN_K4_Lightning(128, L, tpl.metadata, batches[0], "adam", lr, regression=False, **kw)
        elif name == "a1-c2-grf":
            m = HGNN_C2_Lightning_Reg(128, L, tpl.metadata, batches[0], "adam", lr, regression=True, grf_dimension=3, **kw)
        else:
            m = COM_HGNN_SYM_Lightning(128, L, tpl.metadata, batches[0], "adam", lr, model_type="heterogeneous_gnn_k4_com", **kw)
        m.load_state_dict(sd)
        return m.to(DEV)

    # (a) autograd path
    mod = make()
    opt = mod.configure_optimizers()
    losses = []
    for i in range(steps):
        opt.zero_grad()
        loss = mod.training_step(batches[i % 2].to(DEV), i)
        loss.backward()
        opt.step()
        losses.append(loss.item())
    for a, b in zip(losses, ref_losses):
        assert abs(a - b) <= 2e-4 * abs(b), (losses, ref_losses)
    # Adam normalises the update (|dp| ~ lr), so compare parameters by absolute distance relative to lr
    worst = max(((p.detach().double().cpu() - ref_params[k[len("model."):]]).abs().max().item(), k)
                for k, p in mod.state_dict().items() if k.startswith("model."))
    assert worst[0] <= 0.05 * lr * steps, worst
    if not cfg.regression:
        assert 0.0 <= mod.acc.item() <= 1.0 and mod.f1_leg0 is not None

    # (b) fused native path
    mod2 = make()
    tr = FusedTrainer(mod2)
    losses2 = [tr.train_step(batches[i % 2].to(DEV)).item() for i in range(steps)]
    for a, b in zip(losses2, ref_losses):
        assert abs(a - b) <= 2e-4 * abs(b), (losses2, ref_losses)
    worst = max(((p.detach().double().cpu() - ref_params[k[len("model."):]]).abs().max().item(), k)
                for k, p in mod2.state_dict().items() if k.startswith("model."))
    assert worst[0] <= 0.05 * lr * steps, worst
    # inference through the trainer equals the module's forward
    with torch.no_grad():
        b = batches[0].to(DEV)
        assert torch.equal(tr.infer(b), mod2.model(b.x_dict, b.edge_index_dict))


def test_metrics_match_reference_literals_on_device():
    """tests/testGnnLightning.py:L466-500 through the device-side metric implementation."""
    from ms_hgnn.lightning_py.gnnLightning import Base_Lightning
    y_pred = torch.tensor([[0.1, 11, 100, 19, 0.12, 0.14, 15, 24.45], [15, 11, 19, 19, 0.9898, 0.14, -10000, 24.45],
                           [0.1, 13, 100, 19, 0.12, -10, 15, -24.45], [15, 11, 200, 19, 0.9898, 0.14, -10000, 44.45],
                           [-0.1, 11, 100, 19, 0.12, 0.14, 15, 24.45], [-15, 11, 19, 19, -0.9898, 0.14, -10000, 24.45],
                           [-0.1, 13, 100, 19, 0.12, -10, 15, -24.45], [-15, 11, 200, 19, -0.9898, 0.14, -10000, 44.45]],
                          dtype=torch.float64, device=DEV)
    y = torch.tensor([[1, 1, 1, 1], [1, 1, 0, 1], [0, 1, 1, 0], [0, 1, 0, 0], [0, 0, 1, 1], [1, 1, 1, 0], [1, 0, 0, 0], [1, 0, 0, 1]],
                     dtype=torch.float64, device=DEV)
    cfg = CONFIGS["mini_cheetah-k4-contact"]
    b = make_batch(cfg, 8, seed=0)
    mod = HGNN_K4_Lightning(128, 1, M.K4_MINI_CHEETAH.metadata, b, regression=False).to(DEV)
    with torch.no_grad():
        mod.step_helper_function(b.to(DEV))          # compiles the engine the loss head belongs to
        mod.calculate_losses_step(y, y_pred)
    des = torch.nn.functional.cross_entropy(y_pred.reshape(32, 2), y.reshape(32).long(), reduction="mean")
    assert abs(mod.ce_loss.item() - des.item()) < 1e-5 * abs(des.item()) + 1e-5
    assert mod.acc.item() == 0.125
    assert abs(mod.f1_leg0.item() - 0.7272727272727272) < 1e-12
    assert mod.f1_leg1.item() == 0.0 and mod.f1_leg2.item() == 0.75 and abs(mod.f1_leg3.item() - 0.8) < 1e-12
    p16, y16 = Base_Lightning.classification_conversion_16_class(None, torch.tensor([[0.9, 0.3, 0.8, 0.55]], dtype=torch.float64),
                                                                 torch.tensor([[1, 0, 1, 1]]))
    assert y16.tolist() == [[11]] and abs(p16[0, 11].item() - 0.2772) < 1e-12 and abs(p16[0, 0].item() - 0.0063) < 1e-12
