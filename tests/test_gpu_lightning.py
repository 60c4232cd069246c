"""GPU tests of the reference-facing layer: Lightning-shaped modules, checkpoints, fused trainer."""
import os

import pytest
import torch

import mshgnn_oracle as O
from helpers import TOL_FP32, oracle_loss, oracle_model, rel_err
from ms_hgnn import _native as N
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


def test_fused_step_metrics_equal_the_op_by_op_metrics():
    """mshgnn_step_metrics (one call per step) vs the torch op-by-op restatement of gnnLightning.py:L131-151 / customMetrics.py
    on random logits: batch values, epoch accumulation over three steps, reset.  Counts exact, CE within 1e-6 relative."""
    from ms_hgnn.lightning_py import customMetrics as CM
    from ms_hgnn.lightning_py.gnnLightning import Base_Lightning
    g = torch.Generator().manual_seed(5)
    fused = CM.FusedStepMetrics()
    ce, acc, f1s = CM.CrossEntropyLossMetric(), CM.MulticlassAccuracy(), [CM.BinaryF1Score() for _ in range(4)]
    helper = Base_Lightning.__new__(Base_Lightning)
    for step, B in enumerate((4096, 777, 16384)):
        y_pred = (torch.randn(B, 8, generator=g) * 2).to(DEV)
        y = (torch.rand(B, 4, generator=g) > 0.4).double().to(DEV)
        if step == 1:
            y_pred[:5, 0] = y_pred[:5, 1]                     # exact per-foot ties: argmax must pick class 0 on both paths
        b = fused.update(N.LOSS_CE2, y_pred.reshape(-1, 2), y, B, 4).clone()
        _, prob, p1 = Base_Lightning.classification_calculate_useful_values(helper, y_pred.double(), B)
        p16, y16 = Base_Lightning.classification_conversion_16_class(helper, p1, y)
        a = acc(torch.argmax(p16, dim=1), y16.squeeze(1))
        p2 = torch.argmax(prob, dim=1).reshape(B, 4)
        f = [f1s[k](p2[:, k], y[:, k]) for k in range(4)]
        c = ce(y_pred.double().reshape(-1, 2), y.reshape(-1).long())
        assert b[2].item() == round(a.item() * B) and abs(b[21].item() - a.item()) < 1e-14     # count exact, ratio to 1 ulp
        for k in range(4):
            assert abs(b[22 + k].item() - f[k].item()) < 1e-12
        assert abs(b[20].item() - c.item()) <= 1e-6 * abs(c.item())
    e = fused.epoch
    assert e[3].item() == 4096 + 777 + 16384 and e[1].item() == 4 * e[3].item()
    assert abs(e[2].item() / e[3].item() - acc.compute().item()) < 1e-14
    for k in range(4):
        assert abs(CM.BinaryF1Score._f1(e[4 + 4 * k], e[5 + 4 * k], e[6 + 4 * k]).item() - f1s[k].compute().item()) < 1e-12
        assert (e[4 + 4 * k:8 + 4 * k].sum().item()) == e[3].item()              # tp + fp + fn + tn = graphs
    assert abs((e[0].float() / e[1]).item() - ce.compute().item()) <= 1e-6 * abs(ce.compute().item())
    fused.reset()
    assert fused.epoch.abs().sum().item() == 0.0
    # regression head
    p = torch.randn(5000, generator=g).to(DEV); t = torch.randn(5000, generator=g, dtype=torch.float64).to(DEV)
    b = fused.update(N.LOSS_MSE, p, t, 5000, 1)
    d = p.double() - t
    assert abs(b[20].item() - (d * d).mean().item()) < 1e-12 and abs(b[21].item() - (d * d).mean().sqrt().item()) < 1e-12
    assert abs(b[22].item() - d.abs().mean().item()) < 1e-12


def test_train_model_and_evaluate_model_rehost(tmp_path, monkeypatch):
    """train_model / evaluate_model (gnnLightning.py:L913-1421) on a device-resident sequence: checkpoint naming and pruning
    policy, resume, and evaluate_model's return tuple recomputed op by op from the checkpoint's own predictions."""
    import numpy as np
    import window_oracle as WO
    from ms_hgnn.lightning_py.gnnLightning import WindowSubset, evaluate_model, train_model
    from ms_hgnn.windows import DeviceSequence, WindowSpec
    monkeypatch.chdir(tmp_path)
    mat = WO.synthetic_mat(1500, seed=21, dtype=np.float32)
    # learnable labels: contact of leg k = sign of a joint velocity channel
    mat["contacts"] = (mat["qd"][:, [2, 5, 8, 11]] > 0).astype(np.float32)
    ds = DeviceSequence(mat, WindowSpec("heterogeneous_gnn_k4", 150, True), DEV, torch.float32)
    n = len(ds)
    tr, va, te = WindowSubset(ds, torch.arange(0, 900)), WindowSubset(ds, torch.arange(900, 1100)), WindowSubset(ds, torch.arange(1100, n))
    kw = dict(normalize=True, disable_logger=True, batch_size=128, num_layers=2, optimizer="adam", lr=1e-3, hidden_size=128,
              regression=False, seed=3, symmetry_mode="MorphSym", group_operator_path=M.cfg_path("mini_cheetah-k4"))
    path = train_model(tr, va, te, epochs=9, **kw)
    ckpts = sorted(os.listdir(path))
    names = [c for c in ckpts if c.endswith(".ckpt")]
    assert "metrics.jsonl" in ckpts and 3 <= len(names) <= 9
    import re
    pat = re.compile(r"^epoch=(\d+)-val_CE_loss=\d+\.\d{5}-val_F1_Score_Leg_Avg=\d+\.\d{5}\.ckpt$")
    epochs_kept = sorted(int(pat.match(c).group(1)) for c in names)
    assert epochs_kept[-3:] == [6, 7, 8]                                   # the three latest epochs are always kept
    import json
    rows = [json.loads(l) for l in open(os.path.join(path, "metrics.jsonl"))]
    val = [r["val_CE_loss"] for r in rows if "val_CE_loss" in r]
    assert len(val) == 9 and all(v == v and v < 1e3 for v in val)
    best7 = sorted(range(9), key=lambda e: val[e])[:7]
    assert set(epochs_kept) == set(best7) | {6, 7, 8}
    last = os.path.join(path, [c for c in names if c.startswith("epoch=8-")][0])
    ck = torch.load(last, weights_only=False)
    assert ck["epoch"] == 8 and ck["global_step"] == 9 * 8 and all(k.startswith("model.") for k in ck["state_dict"])
    assert 0 < len(ck["optimizer_states"][0]["state"]) <= len(ck["state_dict"])
    # resume: two more epochs continue the step count and the Adam moments
    path2 = train_model(tr, va, te, epochs=11, ckpt_path=last, **kw)
    rows2 = [json.loads(l) for l in open(os.path.join(path2, "metrics.jsonl")) if "epoch" in l]
    assert [r["epoch"] for r in rows2 if "val_CE_loss" in r] == [9, 10] and rows2[-1]["global_step"] == 11 * 8
    # evaluate_model
    pred, lab, acc, f0, f1, f2, f3, favg = evaluate_model(last, te, symmetry_mode="MorphSym", group_operator_path=M.cfg_path("mini_cheetah-k4"),
                                                          batch_size=100, task_type="classification")
    assert pred.shape == lab.shape == (len(te),)
    assert abs(acc.item() - (pred == lab).double().mean().item()) < 1e-12
    bits = lambda v, k: (v >> (3 - k)) & 1
    for k, f in enumerate((f0, f1, f2, f3)):
        p, t = bits(pred, k).bool(), bits(lab, k).bool()
        tp, fp, fn = (p & t).sum().double(), (p & ~t).sum().double(), (~p & t).sum().double()
        ref = torch.nan_to_num(2 * (tp / (tp + fp)) * (tp / (tp + fn)) / (tp / (tp + fp) + tp / (tp + fn)))
        assert abs(f.item() - ref.item()) < 1e-12
    assert abs(favg.item() - (f0 + f1 + f2 + f3).item() / 4) < 1e-12
    # the optimiser made progress on the data it saw: accuracy on the training windows far above the 1/16 chance level
    acc_tr = evaluate_model(last, tr, symmetry_mode="MorphSym", group_operator_path=M.cfg_path("mini_cheetah-k4"), batch_size=300,
                            task_type="classification")[2]
    assert acc_tr.item() > 0.5
