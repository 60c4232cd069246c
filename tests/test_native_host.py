"""CPU checks of the C-ABI library and the host-side plan compiler (no compute calls)."""
import ctypes
import os
import re

import pytest

from ms_hgnn import _native as N
from ms_hgnn import morphology as M
from ms_hgnn.engine import build_spec

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    lib = N.lib()
    header = open(os.path.join(ROOT, "include", "mshgnn_b200.h")).read()
    declared = set(re.findall(r"\b(mshgnn_[a-z_0-9]+)\s*\(", header))
    assert declared, "no declarations found"
    assert declared == set(N.EXPORTS)
    for name in declared:
        assert hasattr(lib, name), name
    assert b"sm_100a" in lib.mshgnn_version()


def _spec(tpl, in_w, morph, dec, C, L=8):
    return build_spec(tpl.node_types, tpl.nodes_per_graph, in_w, tpl.edge_types, tpl.edges, ("gt", "gs", "center_bb"),
                      128, L, morph, "base" if morph else None, dec, C, {}, None)


def test_param_counts_match_reference_models():
    # parameter counts of the BASELINE configs
    k4 = N.NativePlan(_spec(M.K4_MINI_CHEETAH, {"base": 900, "joint": 300, "foot": 900}, True, "foot", 2))
    assert k4.n_params == 2144642
    c2 = N.NativePlan(_spec(M.C2_MINI_CHEETAH, {"base": 900, "joint": 300, "foot": 900}, True, "foot", 2))
    assert c2.n_params == 2407810
    a1 = N.NativePlan(_spec(M.C2_A1, {"base": 900, "joint": 450, "foot": 1}, True, "foot", 3))
    assert a1.n_params == 2312067
    com = N.NativePlan(_spec(M.K4_SOLO_COM, {"base": 6, "joint": 2}, True, "base", 6))
    assert com.n_params == 1350918
    # the two shipped MI-HGNN checkpoints
    mi = N.NativePlan(_spec(M.MI_QUADRUPED, {"base": 6, "joint": 3, "foot": 1}, False, "foot", 1))
    assert mi.n_params == 1317633
    mi2 = N.NativePlan(_spec(M.MI_QUADRUPED, {"base": 900, "joint": 300, "foot": 900}, False, "foot", 2))
    assert mi2.n_params == 1585282


def test_live_row_gemv_counts():
    """Live 128x128 row-GEMVs per graph per layer from the compiled gather lists.

    SURVEY 8d counts 64 (K4) / 52 (C2, K4-COM) per full layer and 8 / 24 in the last one.  The compiled lists
    issue one chunk per (destination node, contributing source), so a thigh (two chain in-edges) costs two
    chunks instead of one pre-summed one: 60 conv + 8 MLP = 68 = 64 + 4.  Liveness is a cone, not just the
    last layer: the decoder reads one node type, so layer L-1 only needs its rows, layer L-2 their
    neighbours, ... (K4: 8, 20, 32, 44 chunks in the last four layers)."""
    k4 = N.NativePlan(_spec(M.K4_MINI_CHEETAH, {"base": 900, "joint": 300, "foot": 900}, True, "foot", 2)).describe()
    assert k4["S"] == 20
    per_layer = [sum(len(t["src"]) for t in layer) for layer in k4["conv"]]
    assert per_layer == [60, 60, 60, 60, 44, 32, 20, 8]
    assert k4["need"][8] == [0] * 16 + [1] * 4                      # decoder reads the feet
    assert k4["need"][7] == [0] * 4 + [0, 0, 1] * 4 + [1] * 4      # calves + feet feed the last layer
    assert k4["need"][6] == [0] * 4 + [0, 1, 1] * 4 + [1] * 4
    assert k4["need"][5] == [0] * 4 + [1] * 12 + [1] * 4
    assert all(k4["need"][l] == [1] * 20 for l in range(5))
    assert k4["mac_rows_fwd"] == 344 + 8 * 4                         # base MLP only where the base rows are live
    c2 = N.NativePlan(_spec(M.C2_MINI_CHEETAH, {"base": 900, "joint": 300, "foot": 900}, True, "foot", 2)).describe()
    assert [sum(len(t["src"]) for t in layer) for layer in c2["conv"]] == [52, 52, 52, 52, 44, 32, 20, 8]
    com = N.NativePlan(_spec(M.K4_SOLO_COM, {"base": 6, "joint": 2}, True, "base", 6)).describe()
    assert [sum(len(t["src"]) for t in layer) for layer in com["conv"]] == [48, 48, 48, 48, 48, 40, 28, 16]
    assert com["need"][8] == [1] * 4 + [0] * 12


def test_weight_gradient_launch_layout():
    """ws_layout (csrc/plan.cu): row splits fitted per layer launch to whole waves of 148 CTAs, partial slots packed."""
    k4 = N.NativePlan(_spec(M.K4_MINI_CHEETAH, {"base": 900, "joint": 300, "foot": 900}, True, "foot", 2))
    before = N.get_option("stack")
    try:
        # stack mode (default): ONE launch for all 94 tasks, one split count fitted to whole waves
        N.set_option("stack", 1)
        for B in (1, 257, 2048, 4096, 16384, 100000):
            lay = k4.dw_layout(B, N.MODE_TC)
            assert len({(ns, rows) for _, ns, rows, _ in lay}) == 1
            tasks, (ns, rows) = sum(t for t, _, _, _ in lay), lay[0][1:3]
            assert tasks == 94 and rows % 64 == 0 and (ns - 1) * rows < B <= ns * rows
            waves = lambda n: -(-tasks * n // 148)
            default_ns = -(-B // 1024)
            assert waves(ns) * (rows + 96) <= waves(default_ns) * (1024 + 96) or B < 1024
            # partial slots of the merged launch are contiguous in task order (layer L-1 first)
            order = sorted(lay, key=lambda e: e[3])
            assert order[0][3] == 0 and all(a[3] + a[0] * a[1] == b[3] for a, b in zip(order, order[1:]))
        assert k4.dw_layout(2048, N.MODE_TC)[0][1:3] == (3, 704)                 # 94 x 3 = 282 CTAs: two waves (was 8 launches of 136 x 256 rows)
        N.set_option("stack", 0)
        _per_layer_layout_checks(k4)
    finally:
        N.set_option("stack", before)


def _per_layer_layout_checks(k4):
    lay = k4.dw_layout(16384, N.MODE_TC)
    assert [t for t, _, _, _ in lay] == [17, 17, 17, 17, 11, 8, 5, 2]          # dead branches prune the last four layers
    assert lay[0][1:3] == (16, 1024)                                             # full layers: the measured default
    assert [(ns, rows) for _, ns, rows, _ in lay[4:]] == [(26, 640), (16, 1024), (29, 576), (64, 256)]
    for B in (1, 96, 257, 1100, 2400, 4096, 16384, 65536, 100000):
        for plan_mode in (N.MODE_TC, N.MODE_FP32):
            lay = k4.dw_layout(B, plan_mode)
            used = []
            for tasks, ns, rows, slot0 in lay:
                assert 1 <= ns <= 64 and B <= ns * rows                             # all rows covered
                if plan_mode == N.MODE_TC:
                    assert rows % 64 == 0 and (ns - 1) * rows < B                   # no empty split
                used.append((slot0, slot0 + tasks * ns))
            used.sort()
            assert used[0][0] == 0 and all(a[1] == b[0] for a, b in zip(used, used[1:]))   # packed, no overlap
    # a pruned launch never takes more waves than the default split would
    for tasks, ns, rows, _ in k4.dw_layout(16384, N.MODE_TC):
        waves = lambda n: -(-tasks * n // 148)
        assert waves(ns) * (rows + 96) <= waves(16) * (1024 + 96)


def test_weight_gradient_launch_layout_any_batch_size():
    """Same invariants for arbitrary batch sizes and the other morphologies (hypothesis)."""
    from hypothesis import given, settings, strategies as st
    plans = [N.NativePlan(_spec(M.K4_MINI_CHEETAH, {"base": 900, "joint": 300, "foot": 900}, True, "foot", 2)),
             N.NativePlan(_spec(M.C2_A1, {"base": 900, "joint": 450, "foot": 1}, True, "foot", 3)),
             N.NativePlan(_spec(M.K4_SOLO_COM, {"base": 6, "joint": 2}, True, "base", 6)),
             N.NativePlan(_spec(M.MI_QUADRUPED, {"base": 6, "joint": 3, "foot": 1}, False, "foot", 1, L=3))]

    @settings(max_examples=200, deadline=None)
    @given(st.integers(0, len(plans) - 1), st.integers(1, 1 << 21), st.sampled_from([N.MODE_TC, N.MODE_TC_1X, N.MODE_FP32]))
    def check(pi, B, mode):
        lay = plans[pi].dw_layout(B, mode)
        used = []
        for tasks, ns, rows, slot0 in lay:
            if tasks == 0:
                continue
            assert 1 <= ns <= 64 and B <= ns * rows
            if mode != N.MODE_FP32:
                assert rows % 64 == 0 and (ns - 1) * rows < B
            used.append((slot0, slot0 + tasks * ns))
        used.sort()
        assert used and used[0][0] == 0 and all(a[1] == b[0] for a, b in zip(used, used[1:]))
        assert plans[pi].workspace_bytes(B, True, mode) > 0

    check()


def test_plan_rejects_unsupported():
    tpl = M.K4_MINI_CHEETAH
    spec = _spec(tpl, {"base": 900, "joint": 300, "foot": 900}, True, "foot", 2)
    spec["hidden"] = 64
    with pytest.raises(RuntimeError, match="H=128"):
        N.NativePlan(spec)
    spec = _spec(tpl, {"base": 900, "joint": 300, "foot": 900}, True, "foot", 2)
    spec["edges"][2] = (1, 1, True, spec["edges"][2][3], spec["edges"][2][4])   # mean over the joint chain: in-degree 2
    with pytest.raises(RuntimeError, match="in-degree"):
        N.NativePlan(spec)


def test_compute_call_without_gpu_fails_loudly():
    import torch
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    plan = N.NativePlan(_spec(M.K4_SOLO_COM, {"base": 6, "joint": 2}, True, "base", 6, L=2))
    buf = ctypes.create_string_buffer(1 << 20)
    addr = (ctypes.addressof(buf) + 255) & ~255
    with pytest.raises(RuntimeError, match="mshgnn_forward failed"):
        plan.forward(1, [addr, addr], N.F32, addr, addr, addr, plan.workspace_bytes(1, False), False, N.MODE_FP32, None)
