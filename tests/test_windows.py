"""Window builder: compiled column tables and the CUDA kernel vs the numpy restatement of the
reference's dataset path (oracle/window_oracle.py).

CPU tests: the z-score pin on the reference's golden matrices, and the compiled tables interpreted in numpy (a test-only
interpreter of the table semantics stated in include/mshgnn_b200.h) against the oracle.  GPU tests: the kernel through
``DeviceSequence`` (ctypes -> C ABI) against the oracle; tolerance 1e-6 absolute on z-scored features for fp64 sequences
(the kernel's statistics are fp64, its outputs fp32), labels exact.
"""
import numpy as np
import pytest
import torch

import window_oracle as WO
from ms_hgnn import morphology as M
from ms_hgnn.windows import DeviceSequence, WindowSpec

T = 150
CASES = [("heterogeneous_gnn_k4", None), ("heterogeneous_gnn_k4", "gs"), ("heterogeneous_gnn_k4", "gt"), ("heterogeneous_gnn_k4", "gr"),
         ("heterogeneous_gnn_c2", None), ("heterogeneous_gnn_c2", "gs"), ("heterogeneous_gnn", None)]


def _kw(model_type, op, mode="MorphSym"):
    if op is None:
        return {}, {}
    name = "mini_cheetah-k4" if model_type.endswith("k4") else "mini_cheetah-c2"
    path = M.cfg_path(name)
    return (dict(symmetry_operator=op, symmetry_mode=mode, group_operator_path=path),
            dict(symmetry_operator=op, symmetry_mode=mode, group=M.load_group(path)))


def interpret_table(spec: WindowSpec, seq: np.ndarray, lab: np.ndarray, starts):
    """Test-only numpy reading of the table contract (mshgnn_window_desc): what the kernel must produce."""
    xs = [[] for _ in spec.nodes]
    ys = []
    for s in starts:
        w = seq[s:s + spec.T].astype(np.float64)
        z = WO.zscore(w) if spec.normalize else w
        b = 0
        for t, (n, k) in enumerate(zip(spec.nodes, spec.blocks)):
            L = spec.block_len[t]
            rows = np.empty((n, k * L))
            for i in range(n):
                for j in range(k):
                    c = spec.block_col[b]
                    rows[i, j * L:(j + 1) * L] = np.asarray(z)[:, c] * spec.block_sign[b] if c >= 0 else float(spec.block_sign[b])
                    b += 1
            xs[t].append(rows)
        ys.append(lab[s + spec.T - 1, spec.label_col] * np.asarray(spec.label_sign))
    return [np.concatenate(v) for v in xs], np.concatenate(ys)


A1_CASES = [(1, False, None), (3, False, None), (3, True, None), (1, True, "gs"), (3, True, "gs")]


def _a1_kw(op):
    if op is None:
        return {}, {}
    path = M.cfg_path("a1-c2")
    return (dict(symmetry_operator=op, symmetry_mode="MorphSym", group_operator_path=path),
            dict(symmetry_operator=op, symmetry_mode="MorphSym", group=M.load_group(path)))


def _a1_oracle(mat, starts, spec_kw, orc_kw, dim, body, normalize=True, T=T):
    parts = [WO.sample_a1(mat, int(i), T, normalize=normalize, grf_dimension=dim, grf_body_to_world_frame=body, **orc_kw) for i in starts]
    xo = [torch.cat([p[k] for p in parts]).numpy() for k in range(3)]
    yo = torch.cat([p[3] for p in parts]).numpy()
    ro = torch.cat([p[4] for p in parts]).numpy() if body else None
    return xo, yo, ro


@pytest.mark.parametrize("dim,body,op", A1_CASES)
def test_a1_tables_match_reference_restatement(dim, body, op):
    """QuadSDKDataset_A1 / heterogeneous_gnn_c2 (BASELINE config a1-c2-grf: widths 900 / 450 / 1, 3-D body-frame GRF labels)."""
    spec_kw, orc_kw = _a1_kw(op)
    mat = WO.synthetic_a1_mat(400, seed=5)
    spec = WindowSpec("heterogeneous_gnn_c2", T, True, dataset="a1", grf_dimension=dim, grf_body_to_world_frame=body, **spec_kw)
    assert spec.widths == {"base": 900, "joint": 450, "foot": 1} and spec.n_labels == (4 if dim == 1 else 12)
    seq, lab = spec.pack(mat, np.float64)
    starts = [0, 7, 250]
    xs, y = interpret_table(spec, seq, lab, starts)
    xo, yo, ro = _a1_oracle(mat, starts, spec_kw, orc_kw, dim, body)
    for k in range(3):
        assert np.abs(xs[k] - xo[k]).max() <= 1e-12
    assert np.abs(y - yo).max() <= 1e-9 * max(1.0, np.abs(yo).max())          # rotation: numpy formula vs scipy
    if body:
        r = xs[3].reshape(len(starts), 4, T)[:, :, -1].reshape(-1)
        assert np.abs(r - ro).max() <= 1e-12


@pytest.mark.parametrize("model_type,n_lab", [("heterogeneous_gnn_k4_com", 24), ("heterogeneous_gnn_c2_com", 12), ("heterogeneous_gnn_s4_com", 6)])
@pytest.mark.parametrize("hist", [1, 3])
def test_solo_tables_match_reference_restatement(model_type, n_lab, hist):
    mat = WO.synthetic_solo_mat(60, seed=2)
    order = [6, 7, 8, 0, 1, 2, 9, 10, 11, 3, 4, 5]
    spec = WindowSpec(model_type, hist, True, dataset="solo12", joint_order=order)
    assert spec.widths == {"base": 6 * hist, "joint": 2 * hist} and spec.n_labels == n_lab
    seq, lab = spec.pack(mat, np.float64)
    starts = [0, 11, 57]
    xs, y = interpret_table(spec, seq, lab, starts)
    parts = [WO.sample_solo(mat, i, model_type, hist, True, order) for i in starts]
    for k in range(2):
        assert np.abs(xs[k] - torch.cat([p[k] for p in parts]).numpy()).max() <= 1e-12
    assert np.abs(y - torch.cat([p[2] for p in parts]).numpy()).max() <= 1e-12
    with pytest.raises(ValueError):
        WindowSpec(model_type, hist, True, dataset="solo12", symmetry_operator="gs")
    with pytest.raises(ValueError):
        WindowSpec("heterogeneous_gnn_k4", hist, True, dataset="solo12")


@pytest.mark.gpu
@pytest.mark.parametrize("dim,body,op", A1_CASES)
def test_a1_kernel_matches_oracle(dim, body, op):
    spec_kw, orc_kw = _a1_kw(op)
    mat = WO.synthetic_a1_mat(500, seed=6)
    spec = WindowSpec("heterogeneous_gnn_c2", T, True, dataset="a1", grf_dimension=dim, grf_body_to_world_frame=body, **spec_kw)
    ds = DeviceSequence(mat, spec, "cuda:0", torch.float64)
    starts = [0, 3, 349, 120, 77]
    b = ds.batch(torch.tensor(starts))
    xo, yo, ro = _a1_oracle(mat, starts, spec_kw, orc_kw, dim, body)
    for k, name in enumerate(("base", "joint", "foot")):
        assert np.abs(b.x_dict[name].cpu().double().numpy() - xo[k]).max() <= 1e-6
    assert np.abs(b.y.cpu().double().numpy() - yo).max() <= 1e-5 * max(1.0, np.abs(yo).max())
    if body:
        assert np.abs(b.r_o.cpu().double().numpy() - ro).max() <= 1e-6
    else:
        assert not hasattr(b, "r_o")


@pytest.mark.gpu
def test_solo_kernel_matches_oracle():
    mat = WO.synthetic_solo_mat(300, seed=3)
    for model_type in ("heterogeneous_gnn_k4_com", "heterogeneous_gnn_c2_com", "heterogeneous_gnn_s4_com"):
        spec = WindowSpec(model_type, 1, True, dataset="solo12")
        ds = DeviceSequence(mat, spec, "cuda:0", torch.float64)
        starts = list(range(0, 300, 7))
        b = ds.batch(torch.tensor(starts))
        parts = [WO.sample_solo(mat, i, model_type, 1, True) for i in starts]
        for k, name in enumerate(("base", "joint")):
            assert np.abs(b.x_dict[name].cpu().double().numpy() - torch.cat([p[k] for p in parts]).numpy()).max() <= 1e-6
        assert np.abs(b.y.cpu().double().numpy() - torch.cat([p[2] for p in parts]).numpy()).max() <= 1e-6


def test_zscore_pinned_by_reference_golden_matrices():
    """tests/testDatasets.py:L513-539 holds z-scored [6, n] matrices (history 6); re-normalising them must be the identity
    iff the standard deviation uses Bessel's correction (flexibleDataset.py:L396)."""
    des_la = np.array([[-0.5362300252378239, -1.6632797193920590, -1.5898042183134058],
                       [0.0423700008727095, -0.2257869376580193, -0.3294529164335535],
                       [0.0423700008727095, -0.2257869376580193, -0.3294529164335535],
                       [-0.0272254120334773, 0.0614106314680039, 0.3784460702629509],
                       [-1.2761523397622021, 1.1253516629229292, 1.3613238439024378],
                       [1.7548677752880986, 0.9280913003171616, 0.5089401370151242]])
    des_av = np.array([[0.3922097071448895, -1.9912456309175439, -1.6041605549730955],
                       [-0.9026040953610488, 0.1920045172187222, -0.2748544676177472],
                       [-0.9026040953610488, 0.1920045172187222, -0.2748544676177472],
                       [-0.7582570552051391, 0.7857144431190286, 0.0388281714139248],
                       [0.7083651482760132, 0.3541357390230436, 0.9208249386706086],
                       [1.4628903905063348, 0.4673864143380300, 1.1942163801239956]])
    for des in (des_la, des_av):
        np.testing.assert_allclose(np.asarray(WO.zscore(des)), des, atol=1e-12)
        biased = (des - des.mean(0)) / des.std(0)          # n instead of n - 1: must NOT reproduce the pins
        assert np.abs(biased - des).max() > 1e-2
    const = np.full((6, 2), 3.25)
    assert np.all(np.asarray(WO.zscore(const)) == 0.0)     # 0/0 -> NaN -> 0


@pytest.mark.parametrize("model_type,op", CASES)
def test_compiled_tables_match_reference_restatement(model_type, op):
    mat = WO.synthetic_mat(400, seed=3)
    skw, okw = _kw(model_type, op)
    spec = WindowSpec(model_type, T, True, **skw)
    seq, lab = spec.pack(mat, np.float64)
    starts = [0, 1, 17, 250]
    xs, y = interpret_table(spec, seq, lab, starts)
    xo, yo = WO.batch(mat, starts, model_type, T, normalize=True, **okw)
    for t, name in enumerate(("base", "joint", "foot")):
        assert xs[t].shape == tuple(xo[name].shape)
        np.testing.assert_allclose(xs[t], xo[name].numpy(), atol=1e-12, rtol=0)
    np.testing.assert_array_equal(y, yo.numpy())


def test_euclidean_mode_and_unnormalised_tables():
    mat = WO.synthetic_mat(300, seed=5)
    skw, okw = _kw("heterogeneous_gnn_k4", "gr", "Euclidean")
    spec = WindowSpec("heterogeneous_gnn_k4", 10, False, **skw)
    seq, lab = spec.pack(mat, np.float64)
    xs, y = interpret_table(spec, seq, lab, [5, 100])
    xo, yo = WO.batch(mat, [5, 100], "heterogeneous_gnn_k4", 10, normalize=False, **okw)
    for t, name in enumerate(("base", "joint", "foot")):
        np.testing.assert_array_equal(xs[t], xo[name].numpy())
    np.testing.assert_array_equal(y, yo.numpy())


def test_argument_errors_mirror_the_reference():
    with pytest.raises(ValueError):
        WindowSpec("heterogeneous_gnn_k4", T, True, symmetry_operator="gs")                      # no mode / path
    with pytest.raises(ValueError):
        WindowSpec("dynamics", T)
    with pytest.raises(RuntimeError, match="no CPU path"):
        DeviceSequence(WO.synthetic_mat(200), WindowSpec("heterogeneous_gnn_k4", T), "cpu")


# ------------------------------------------------------------------------------------------ GPU
@pytest.mark.gpu
@pytest.mark.parametrize("model_type,op", CASES)
@pytest.mark.parametrize("dtype", [torch.float64, torch.float32])
def test_kernel_matches_oracle(model_type, op, dtype):
    mat = WO.synthetic_mat(600, seed=11, dtype=np.float64 if dtype == torch.float64 else np.float32)
    skw, okw = _kw(model_type, op)
    spec = WindowSpec(model_type, T, True, **skw)
    ds = DeviceSequence(mat, spec, "cuda:0", dtype)
    assert len(ds) == 600 - T + 1
    g = torch.Generator().manual_seed(1)
    idx = torch.cat((torch.tensor([0, len(ds) - 1, 7, 7]), torch.randint(0, len(ds), (61,), generator=g)))
    b = ds.batch(idx)
    xo, yo = WO.batch(mat, idx.tolist(), model_type, T, normalize=True, **okw)
    # fp64 sequence: fp64 statistics, fp32 rounding of the result only.  fp32 sequence: the oracle sees the same fp32
    # samples (promoted exactly to fp64), so the same bound holds.
    for name in ("base", "joint", "foot"):
        got = b.x_dict[name].cpu().double()
        assert got.shape == xo[name].shape
        assert (got - xo[name]).abs().max().item() <= 1e-6
    assert torch.equal(b.y.cpu().double(), yo)
    ei = b.edge_index_dict
    for et, v in spec.template.edge_index_dict(idx.numel()).items():
        assert torch.equal(ei[et].cpu(), v)                      # bit-exact batching


@pytest.mark.gpu
def test_kernel_unpadded_rows_take_the_element_load_path():
    """54-column rows are not 16-byte multiples: the kernel falls back from the bulk async copy to element loads."""
    mat = WO.synthetic_mat(500, seed=13)
    spec = WindowSpec("heterogeneous_gnn_k4", T, True)
    spec.seq_cols = 54
    ds = DeviceSequence(mat, spec, "cuda:0", torch.float64)
    assert ds.seq.shape[1] == 54
    idx = torch.arange(0, len(ds), 3)
    b = ds.batch(idx)
    xo, yo = WO.batch(mat, idx.tolist(), "heterogeneous_gnn_k4", T, normalize=True)
    for name in ("base", "joint", "foot"):
        assert (b.x_dict[name].cpu().double() - xo[name]).abs().max().item() <= 1e-6
    assert torch.equal(b.y.cpu().double(), yo)


@pytest.mark.gpu
def test_kernel_unnormalised_short_history_and_reuse():
    mat = WO.synthetic_mat(64, seed=2)
    spec = WindowSpec("heterogeneous_gnn_c2", 5, False)
    ds = DeviceSequence(mat, spec, "cuda:0", torch.float64)
    idx = torch.arange(len(ds))
    b = ds.batch(idx)
    xo, yo = WO.batch(mat, idx.tolist(), "heterogeneous_gnn_c2", 5, normalize=False)
    for name in ("base", "joint", "foot"):
        assert torch.equal(b.x_dict[name].cpu(), xo[name].float())
    b2 = ds.batch(idx.flip(0).cuda(), out=b)                      # device indices, buffers reused
    assert b2 is b
    assert torch.equal(b.y.cpu().double(), yo.reshape(-1, 4).flip(0).reshape(-1))
    with pytest.raises(IndexError):
        ds.batch(torch.tensor([len(ds)]))


@pytest.mark.gpu
def test_windowed_batch_feeds_the_model_like_a_host_batch():
    """Full-size property (16384 graphs): a batch built on the device equals the host-collated one bit for bit after the
    same fp32 rounding, so the native forward gives identical logits for both."""
    from ms_hgnn.synthetic import CONFIGS, HeteroBatch, build_model
    mat = WO.synthetic_mat(16384 + T - 1, seed=4, dtype=np.float32)
    spec = WindowSpec("heterogeneous_gnn_k4", T, True)
    ds = DeviceSequence(mat, spec, "cuda:0", torch.float32)
    idx = torch.randperm(len(ds), generator=torch.Generator().manual_seed(9))
    b = ds.batch(idx)
    assert b.batch_size == 16384
    # z-scored blocks: mean 0, Bessel std 1 (or all-zero for the constant column) for EVERY block of every graph
    for name, k in (("base", 6), ("joint", 2), ("foot", 6)):
        v = b.x_dict[name].view(-1, k, T).double()
        assert v.mean(-1).abs().max().item() < 1e-5
        sd = v.std(-1)
        assert ((sd - 1).abs() < 1e-5).logical_or(sd == 0).all()
    # spot-check 32 graphs against the oracle and run the model on both
    sel = torch.arange(0, 16384, 512)
    xo, yo = WO.batch(mat, idx[sel].tolist(), "heterogeneous_gnn_k4", T, normalize=True)
    for name, n in (("base", 4), ("joint", 12), ("foot", 4)):
        rows = (sel[:, None] * n + torch.arange(n)[None]).reshape(-1)
        assert (b.x_dict[name][rows.cuda()].cpu().double() - xo[name]).abs().max().item() <= 1e-6
    cfg = CONFIGS["mini_cheetah-k4-contact"]
    nm = build_model(cfg, layers=8, seed=3).to("cuda:0")
    with torch.no_grad():
        y_dev = nm(b.x_dict, b.edge_index_dict)
        host = HeteroBatch({k: v.cpu() for k, v in b.x_dict.items()}, {k: v.cpu() for k, v in b.edge_index_dict.items()}, b.y.cpu(), 16384)
        hb = host.to("cuda:0")
        y_host = nm(hb.x_dict, hb.edge_index_dict)
    assert torch.equal(y_dev, y_host)


@pytest.mark.gpu
@pytest.mark.parametrize("cfg_name,dataset,model_type,hist,kw", [
    ("a1-c2-grf", "a1", "heterogeneous_gnn_c2", T, dict(grf_dimension=3, grf_body_to_world_frame=True)),
    ("solo12-k4-com", "solo12", "heterogeneous_gnn_k4_com", 1, {}),
    ("solo-c2-com", "solo12", "heterogeneous_gnn_c2_com", 1, {})])
def test_other_datasets_feed_their_models(cfg_name, dataset, model_type, hist, kw):
    """The A1 and Solo12 window layouts line up with the models of the BASELINE configs that consume them (feature widths, label
    count, edge template): one native train step on a windowed batch, loss equal to the oracle model's on the same tensors."""
    from helpers import TOL_FP32, oracle_model, oracle_run
    from ms_hgnn import _native as N
    from ms_hgnn.synthetic import CONFIGS, HeteroBatch, build_model
    cfg = CONFIGS[cfg_name]
    mat = WO.synthetic_a1_mat(600, seed=8, dtype=np.float32) if dataset == "a1" else WO.synthetic_solo_mat(600, seed=8)
    spec = WindowSpec(model_type, hist, True, dataset=dataset, **kw)
    assert spec.widths == cfg.in_width and spec.n_labels == cfg.label_width
    ds = DeviceSequence(mat, spec, "cuda:0", torch.float32)
    idx = torch.randperm(len(ds), generator=torch.Generator().manual_seed(3))[:96]
    b = ds.batch(idx)
    om = oracle_model(cfg, layers=4, seed=1)
    nm = build_model(cfg, layers=4, seed=2)
    nm.load_state_dict({k: v.float() for k, v in om.state_dict().items()})
    nm = nm.to("cuda:0")
    out = nm(b.x_dict, b.edge_index_dict)
    eng = nm._last_engine
    loss, dout = eng.loss(out.detach().reshape(-1, eng.spec["out_channels"]).float().contiguous(), b.y, N.LOSS_MSE)
    out.backward(dout.view_as(out).to(out.dtype))
    host = HeteroBatch({k: v.cpu() for k, v in b.x_dict.items()}, {k: v.cpu() for k, v in b.edge_index_dict.items()}, b.y.cpu(), 96)
    out_o, loss_o, _ = oracle_run(cfg, om, host)
    assert abs(loss.item() - loss_o.item()) <= TOL_FP32 * abs(loss_o.item())
    assert all(torch.isfinite(p.grad).all() for p in nm.parameters() if p.grad is not None)
