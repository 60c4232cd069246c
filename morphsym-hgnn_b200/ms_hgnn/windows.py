"""Device-side dataset: sliding windows of a raw sequence -> collated HeteroData node features, on the GPU.

The reference materialises every 150-row window on the host, per sample, in fp64
(``LinTzuYaunDataset.load_data_at_dataset_seq`` LinTzuYaunDataset.py:L66-88, ``load_data_sorted_k4/_c2``
LinTzuYaunDataset_Morph.py:L156-347, ``get_helper_heterogeneous_gnn(_c2)`` L555-697, ``FlexibleDataset.get``
flexibleDataset.py:L444-470) and the DataLoader collates them: 43.2 KB of features per graph cross the host link
although consecutive windows share 149 of their 150 rows.  Here the raw sequence (216 B per time step for the Mini
Cheetah data) is uploaded once; a batch is a tensor of window start rows, and ``mshgnn_build_windows`` writes the
collated ``x_dict`` / ``y`` straight into device memory.  The column order (URDF sort), the base tiling
and the optional dataset-level group action (``apply_symmetry`` L349-408) are compiled once into a
column / sign table; the kernel reads no index tensor besides the start rows.

There is no CPU path: ``DeviceSequence`` requires a CUDA device and raises otherwise.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, List, Optional, Sequence

import numpy as np
import torch

from . import _native as N
from . import morphology as M
from .synthetic import HeteroBatch

# raw channel arrays of data.mat in the order they are packed side by side (LinTzuYaunDataset.py:L79-85; tau_est is
# loaded by the reference but returned as None, so it never reaches the model)
MINI_CHEETAH_CHANNELS = (("imu_acc", 3), ("imu_omega", 3), ("q", 12), ("qd", 12), ("p", 12), ("v", 12))
# LinTzuYaunDataset.get_urdf_name_to_dataset_array_index (L34-60) composed with the URDF node order RL, FL, RR, FR
# (flexibleDataset.py:L131-142)
MINI_CHEETAH_JOINT_ORDER = (9, 10, 11, 3, 4, 5, 6, 7, 8, 0, 1, 2)
MINI_CHEETAH_FOOT_ORDER = (3, 1, 2, 0)

_MODEL_TYPES = {"heterogeneous_gnn": ("mi_quadruped", 1), "heterogeneous_gnn_k4": ("k4_mini_cheetah", 4),
                "heterogeneous_gnn_c2": ("c2_mini_cheetah", 2)}

# QuadSDKDataset (A1, Gazebo): data.mat arrays the model path reads (quadSDKDataset.py:L107-117, quadSDKDataset_Morph.py:L464-482).
# URDF node order of a1_pruned.urdf = dataset order (joints '8','0','1' | '9','2','3' | '10','4','5' | '11','6','7', toes jtoe0..3:
# quadSDKDataset_Morph.py:L404-441), so joint_node_indices_sorted / foot_node_indices_sorted are identities.
A1_CHANNELS = (("imu_acc", 3), ("imu_omega", 3), ("q", 12), ("qd", 12), ("tau", 12))
# Solo12 centre-of-mass data: X = [q (12) | qd (12)], Y = [base lin vel (3) | base ang vel (3)] (soloDataset.py:L382-401)
SOLO_CHANNELS = (("q", 12), ("qd", 12))
_SOLO_MODEL_TYPES = {"heterogeneous_gnn_k4_com": ("k4_solo_com", 4), "heterogeneous_gnn_c2_com": ("c2_solo_com", 2),
                     "heterogeneous_gnn_s4_com": ("s4_solo_com", 1)}


def _compose(perm: Sequence[Sequence[int]], coef: Sequence[Sequence[float]], op: Optional[str], n: int, morphsym: bool):
    """(source index, factor) of every output column j of ``apply_symmetry``: out[:, j] = in[:, src[j]] * f[j]."""
    if op is None:
        return list(range(n)), [1] * n
    one = [1] * n
    gs = [int(v) for v in coef[0]] if morphsym else one
    gt = [int(v) for v in coef[1]] if morphsym else one
    if op == "gs":
        return [int(perm[0][j]) for j in range(n)], gs
    if op == "gt":
        return [int(perm[1][j]) for j in range(n)], gt
    if op == "gr":      # data[:, P0][:, P1] * (gs * gt)
        return [int(perm[0][int(perm[1][j])]) for j in range(n)], [gs[j] * gt[j] for j in range(n)]
    raise ValueError(f"symmetry_operator must be 'gs', 'gt', 'gr' or None, not {op!r}")


def quat_rotate_rows(quat_xyzw: np.ndarray, vec: np.ndarray) -> np.ndarray:
    """``Rotation.from_quat(q).as_matrix() @ v`` for every row (scipy normalises the quaternion first):
    the per-entry label rotation of quadSDKDataset_Morph.py:L476-480, done once for the whole sequence."""
    q = np.asarray(quat_xyzw, dtype=np.float64)
    q = q / np.linalg.norm(q, axis=1, keepdims=True)
    x, y, z, w = q[:, 0], q[:, 1], q[:, 2], q[:, 3]
    R = np.empty((q.shape[0], 3, 3))
    R[:, 0, 0] = 1 - 2 * (y * y + z * z); R[:, 0, 1] = 2 * (x * y - z * w); R[:, 0, 2] = 2 * (x * z + y * w)
    R[:, 1, 0] = 2 * (x * y + z * w); R[:, 1, 1] = 1 - 2 * (x * x + z * z); R[:, 1, 2] = 2 * (y * z - x * w)
    R[:, 2, 0] = 2 * (x * z - y * w); R[:, 2, 1] = 2 * (y * z + x * w); R[:, 2, 2] = 1 - 2 * (x * x + y * y)
    return np.einsum("nij,nkj->nki", R, np.asarray(vec, dtype=np.float64))


class WindowSpec:
    """Compiled column / sign tables of one (dataset, model_type, symmetry) combination.

    ``dataset``: ``"mini_cheetah"`` (LinTzuYaunDataset_Morph, contact labels), ``"a1"`` (QuadSDKDataset_NewGraph / QuadSDKDataset_A1:
    C2 graph, joint torques as a third joint variable, constant foot feature, 1-D or 3-D ground-reaction-force labels, optionally
    rotated by the body orientation) or ``"solo12"`` (Solo12Dataset: centre-of-mass momentum regression, zero base features,
    dataset-level standardisation instead of the per-window z-score)."""

    def __init__(self, model_type: str, history_length: int, normalize: bool = True, symmetry_operator: Optional[str] = None,
                 symmetry_mode: Optional[str] = None, group_operator_path: Optional[str] = None, dataset: str = "mini_cheetah",
                 grf_dimension: int = 1, grf_body_to_world_frame: bool = False, joint_order: Optional[Sequence[int]] = None):
        if dataset not in ("mini_cheetah", "a1", "solo12"):
            raise ValueError(f"dataset {dataset!r}: 'mini_cheetah', 'a1' or 'solo12'")
        self.dataset = dataset
        self.extra_types: List[str] = []          # pseudo node types of the kernel launch that are not model inputs ("r_o")
        self.grf_dimension, self.grf_body_to_world_frame = int(grf_dimension), bool(grf_body_to_world_frame)
        if dataset == "a1":
            self._init_a1(model_type, history_length, normalize, symmetry_operator, symmetry_mode, group_operator_path)
            return
        if dataset == "solo12":
            self._init_solo(model_type, history_length, normalize, symmetry_operator, joint_order)
            return
        if model_type not in _MODEL_TYPES:
            raise ValueError(f"model_type {model_type!r} has no heterogeneous graph layout")
        # same argument checks as LinTzuYaunDataset_Morph.__init__ (L40-46)
        if symmetry_operator is not None and ((symmetry_mode != "MorphSym" and symmetry_mode != "Euclidean") or group_operator_path is None):
            raise ValueError("symmetry_mode must be 'MorphSym' or 'Euclidean' when symmetry_operator is not None.")
        if symmetry_operator is not None and model_type == "heterogeneous_gnn":
            raise ValueError("the dataset-level group action exists only for the k4 / c2 graph layouts")
        if normalize and history_length < 2:
            raise ValueError("normalize=True needs history_length >= 2")
        tpl_name, nb = _MODEL_TYPES[model_type]
        self.template = M.TEMPLATES[tpl_name]
        self.model_type, self.T, self.normalize = model_type, int(history_length), bool(normalize)
        self.channels = MINI_CHEETAH_CHANNELS
        off, o = {}, 0
        for name, w in MINI_CHEETAH_CHANNELS:
            off[name] = o
            o += w
        self.seq_cols = (o + 3) // 4 * 4        # rows padded to a 16-byte multiple: one bulk async copy per window
        group = M.load_group(group_operator_path) if symmetry_operator is not None else None
        ms = symmetry_mode == "MorphSym"
        op = symmetry_operator
        g = group or {}
        # base: np.tile(raw, (1, nb)) then apply_symmetry('base'); tiled column m reads raw column m % 3
        bsrc, blin = _compose(g.get("permutation_Q_bs"), g.get("reflection_Q_bs_lin"), op, 3 * nb, ms)
        _, bang = _compose(g.get("permutation_Q_bs"), g.get("reflection_Q_bs_ang"), op, 3 * nb, ms)
        jsrc, jf = _compose(g.get("permutation_Q_js"), g.get("reflection_Q_js"), op, 12, ms)
        fsrc, ff = _compose(g.get("permutation_Q_fs"), g.get("reflection_Q_fs"), op, 12, ms)
        lsrc, lf = _compose(g.get("permutation_Q_ls"), g.get("reflection_Q_ls"), op, 4, ms)
        col: List[int] = []
        sign: List[int] = []
        for i in range(nb):                                 # base node i: [lin_acc xyz | ang_vel xyz], each flatten('F')
            for name, f in (("imu_acc", blin), ("imu_omega", bang)):
                for a in range(3):
                    col.append(off[name] + bsrc[3 * i + a] % 3)
                    sign.append(f[3 * i + a])
        for i in range(12):                                 # joint node i: [q | qd]
            for name in ("q", "qd"):
                col.append(off[name] + MINI_CHEETAH_JOINT_ORDER[jsrc[i]])
                sign.append(jf[i])
        for i in range(4):                                  # foot node i: [p xyz | v xyz]
            for name in ("p", "v"):
                for a in range(3):
                    m = fsrc[3 * i + a]                     # column of the URDF-sorted foot array
                    col.append(off[name] + 3 * MINI_CHEETAH_FOOT_ORDER[m // 3] + m % 3)
                    sign.append(ff[3 * i + a])
        self.nodes = [nb, 12, 4]
        self.blocks = [6, 2, 6]
        self.block_len = [self.T] * 3
        self.block_col, self.block_sign = col, sign
        self.label_col = [MINI_CHEETAH_FOOT_ORDER[lsrc[i]] for i in range(4)]
        self.label_sign = list(lf)
        self.label_cols = 4

    def _init_a1(self, model_type, history_length, normalize, op, symmetry_mode, group_operator_path):
        """QuadSDKDataset_NewGraph.load_data_sorted_c2 / get_helper_heterogeneous_gnn_c2 (quadSDKDataset_Morph.py:L100-174, L274-351)."""
        if model_type != "heterogeneous_gnn_c2":
            raise ValueError(f"Invalid model type: {model_type}")                         # quadSDKDataset_Morph.py:L82-85
        if self.grf_dimension not in (1, 3):
            raise ValueError(f"Invalid grf_dimension: {self.grf_dimension}")
        if op is not None and ((symmetry_mode != "MorphSym" and symmetry_mode != "Euclidean") or group_operator_path is None):
            raise ValueError("symmetry_mode must be 'MorphSym' or 'Euclidean' when symmetry_operator is not None.")
        if normalize and history_length < 2:
            raise ValueError("normalize=True needs history_length >= 2")
        self.template = M.TEMPLATES["c2_a1"]
        self.model_type, self.T, self.normalize = model_type, int(history_length), bool(normalize)
        self.channels = A1_CHANNELS + ((("r_o", 4),) if self.grf_body_to_world_frame else ())
        off, o = {}, 0
        for name, w in self.channels:
            off[name] = o
            o += w
        self.seq_cols = (o + 3) // 4 * 4
        g = M.load_group(group_operator_path) if op is not None else {}
        ms = symmetry_mode == "MorphSym"
        nb = 2
        bsrc, blin = _compose(g.get("permutation_Q_bs"), g.get("reflection_Q_bs_lin"), op, 3 * nb, ms)
        _, bang = _compose(g.get("permutation_Q_bs"), g.get("reflection_Q_bs_ang"), op, 3 * nb, ms)
        jsrc, jf = _compose(g.get("permutation_Q_js"), g.get("reflection_Q_js"), op, 12, ms)
        col: List[int] = []
        sign: List[int] = []
        for i in range(nb):
            for name, f in (("imu_acc", blin), ("imu_omega", bang)):
                for a in range(3):
                    col.append(off[name] + bsrc[3 * i + a] % 3)
                    sign.append(f[3 * i + a])
        for i in range(12):                                 # joint node i: [q | qd | tau]
            for name in ("q", "qd", "tau"):
                col.append(off[name] + jsrc[i])
                sign.append(jf[i])
        for i in range(4):                                  # no foot variables: torch.ones((4, foot_width = 1)) (flexibleDataset.py:L187-190, L565)
            col.append(-1)
            sign.append(1)
        self.nodes, self.blocks, self.block_len = [nb, 12, 4], [6, 3, 1], [self.T, self.T, 1]
        if self.grf_body_to_world_frame:                    # data.r_o = r_o[-1] of the (z-scored when normalize) quaternion history (L346-348)
            for a in range(4):
                col.append(off["r_o"] + a)
                sign.append(1)
            self.nodes.append(1); self.blocks.append(4); self.block_len.append(self.T)
            self.extra_types = ["r_o"]
        self.block_col, self.block_sign = col, sign
        if self.grf_dimension == 1:                         # z components, apply_symmetry(part='label') with Q_ls (L147-149, L199-202)
            lsrc, lf = _compose(g.get("permutation_Q_ls"), g.get("reflection_Q_ls"), op, 4, ms)
            self.label_col = [3 * lsrc[i] + 2 for i in range(4)]
        else:                                               # 3-D: the foot-space representation Q_fs (L150-156, L203-206)
            lsrc, lf = _compose(g.get("permutation_Q_fs"), g.get("reflection_Q_fs"), op, 12, ms)
            self.label_col = [lsrc[i] for i in range(12)]
        self.label_sign = list(lf)
        self.label_cols = 12

    def _init_solo(self, model_type, history_length, normalize, op, joint_order):
        """Solo12Dataset.load_data_sorted(_k4 / _c2) / get_helper_heterogeneous_gnn (soloDataset.py:L235-300, L332-401, L546-718)."""
        if model_type not in _SOLO_MODEL_TYPES:
            raise ValueError(f"Model type {model_type} is not implemented for Solo12 dataset.")
        if op is not None:
            # the reference's label branch of apply_symmetry indexes a python list with a list (soloDataset.py:L611-612 -> L733) and
            # reflection_Q_ls is never loaded (L103): the dataset-level group action cannot run there, so it is not offered here
            raise ValueError("Solo12Dataset: symmetry_operator is not usable in the reference (soloDataset.py:L611, L733)")
        tpl_name, nb = _SOLO_MODEL_TYPES[model_type]
        self.template = M.TEMPLATES[tpl_name]
        # normalize = dataset-level standardisation with stored statistics (L136-143), applied once in pack(); no per-window z-score
        self.model_type, self.T, self.normalize, self.standardize = model_type, int(history_length), False, bool(normalize)
        self.channels = SOLO_CHANNELS
        self.seq_cols = 24
        # joint_node_indices_sorted (flexibleDataset.py:L138-142) depends on urdf_files/Solo/solo12.urdf, which the reference tree does
        # not contain: identity (dataset order FL, FR, HL, HR) unless the caller passes the URDF node order
        order = list(range(12)) if joint_order is None else [int(v) for v in joint_order]
        if sorted(order) != list(range(12)):
            raise ValueError("joint_order must be a permutation of 0..11")
        self.joint_order = order
        col: List[int] = []
        sign: List[int] = []
        for i in range(nb):                                 # lin_vel / ang_vel inputs are zeros (L396-397), tiled per base node
            for _ in range(6):
                col.append(-1)
                sign.append(0)
        for i in range(12):
            for base in (0, 12):
                col.append(base + order[i])
                sign.append(1)
        self.nodes, self.blocks, self.block_len = [nb, 12], [6, 2], [self.T, self.T]
        self.block_col, self.block_sign = col, sign
        # labels [lin(3) | ang(3)] of the last frame, repeated per base node (L605-618, L706-716; one copy for S4: L380)
        self.label_col = [c for _ in range(nb) for c in range(6)]
        self.label_sign = [1] * (6 * nb)
        self.label_cols = 6

    @property
    def n_labels(self) -> int:
        return len(self.label_col)

    @property
    def kernel_types(self) -> List[str]:
        return list(self.template.node_types) + list(self.extra_types)

    @property
    def widths(self) -> Dict[str, int]:
        return {t: self.blocks[k] * self.block_len[k] for k, t in enumerate(self.template.node_types)}

    def pack(self, mat: Dict[str, np.ndarray], dtype=np.float32):
        """Raw arrays -> (seq [n_rows, seq_cols] (channels side by side + zero padding), labels [n_rows, label_cols])."""
        if self.dataset == "solo12":
            X, Y = np.asarray(mat["X"], dtype=np.float64), np.asarray(mat["Y"], dtype=np.float64)
            if self.standardize:                            # Standarizer.transform (soloDataset.py:L18-31) with rss_stats.npz
                X = (X - np.asarray(mat["x_mean"])) / np.asarray(mat["x_std"])
                Y = (Y - np.asarray(mat["y_mean"])) / np.asarray(mat["y_std"])
            return np.ascontiguousarray(X.astype(dtype)), np.ascontiguousarray(Y.astype(dtype))
        lab_key = "F" if self.dataset == "a1" else "contacts"
        n = int(np.asarray(mat[lab_key]).shape[0])
        seq = np.concatenate([np.asarray(mat[name]).reshape(n, w) for name, w in self.channels], axis=1).astype(dtype)
        if seq.shape[1] < self.seq_cols:
            seq = np.concatenate([seq, np.zeros((n, self.seq_cols - seq.shape[1]), dtype=dtype)], axis=1)
        lab = np.asarray(mat[lab_key]).reshape(n, self.label_cols).astype(np.float64)
        if self.dataset == "a1" and self.grf_body_to_world_frame:
            lab = quat_rotate_rows(np.asarray(mat["r_o"]).reshape(n, 4), lab.reshape(n, 4, 3)).reshape(n, 12)
        return np.ascontiguousarray(seq), np.ascontiguousarray(lab.astype(dtype))


class DeviceSequence:
    """One dataset sequence resident on the GPU; ``batch(idx)`` is ``Batch.from_data_list([dataset[i] for i in idx])``.

    ``len()`` and index semantics follow ``FlexibleDataset`` (flexibleDataset.py:L90: ``length = entries - history_length + 1``;
    entry ``i`` is the window of rows ``[i, i + history_length)`` labelled by its last row)."""

    def __init__(self, mat: Dict[str, np.ndarray], spec: WindowSpec, device, dtype=torch.float32):
        device = torch.device(device)
        if device.type != "cuda":
            raise RuntimeError("DeviceSequence builds windows with the native CUDA library: no CPU path")
        if dtype not in (torch.float32, torch.float64):
            raise ValueError("dtype must be float32 or float64")
        self.spec, self.device, self.dtype = spec, device, dtype
        seq, lab = spec.pack(mat, np.float64 if dtype == torch.float64 else np.float32)
        self.n_rows = seq.shape[0]
        if lab.shape[0] != self.n_rows:
            raise ValueError("features and labels have different numbers of rows")
        if self.n_rows < spec.T:
            raise ValueError("Dataset has too few entries for the provided 'history_length'.")
        self.seq = torch.from_numpy(seq).to(device)
        self.labels = torch.from_numpy(lab).to(device)
        self._edges: Dict[int, dict] = {}
        s = spec
        n = len(s.block_col)
        nl = s.n_labels
        nt = len(s.nodes)
        self._keep = [(C.c_int32 * n)(*s.block_col), (C.c_int32 * n)(*s.block_sign), (C.c_int32 * nl)(*s.label_col), (C.c_int32 * nl)(*s.label_sign)]
        d = N.WindowDesc()
        d.history_length, d.seq_cols, d.label_cols, d.n_node_types = s.T, s.seq_cols, s.label_cols, nt
        for t in range(nt):
            d.nodes_per_graph[t], d.blocks_per_node[t], d.block_len[t] = s.nodes[t], s.blocks[t], s.block_len[t]
        d.normalize, d.n_labels = int(s.normalize), nl
        d.block_col = C.cast(self._keep[0], C.POINTER(C.c_int32)); d.block_sign = C.cast(self._keep[1], C.POINTER(C.c_int32))
        d.label_col = C.cast(self._keep[2], C.POINTER(C.c_int32)); d.label_sign = C.cast(self._keep[3], C.POINTER(C.c_int32))
        self._desc = d

    def __len__(self) -> int:
        return self.n_rows - self.spec.T + 1

    def get_data_metadata(self):
        return self.spec.template.metadata

    def get_data_format(self) -> str:
        return self.spec.model_type

    def edge_index_dict(self, B: int):
        if B not in self._edges:
            self._edges[B] = self.spec.template.edge_index_dict(B, self.device)
        return self._edges[B]

    def batch(self, idx: torch.Tensor, out: Optional[HeteroBatch] = None) -> HeteroBatch:
        """Collated batch of the dataset entries ``idx`` (int64; host or device).  ``out``: a batch of the same size
        returned earlier, whose tensors are overwritten instead of allocating new ones."""
        idx = torch.as_tensor(idx, dtype=torch.int64)
        if idx.ndim != 1 or idx.numel() < 1:
            raise ValueError("idx must be a non-empty 1-D index tensor")
        if idx.device.type != "cuda":
            if int(idx.min()) < 0 or int(idx.max()) >= len(self):
                raise IndexError("dataset index out of range")
            idx = idx.to(self.device, non_blocking=True)
        B = idx.numel()
        s = self.spec
        names = s.template.node_types
        kt = s.kernel_types
        nl = s.n_labels
        if out is None:
            x = {t: torch.empty(B * s.nodes[k], s.blocks[k] * s.block_len[k], dtype=torch.float32, device=self.device) for k, t in enumerate(names)}
            y = torch.empty(B * nl, dtype=torch.float32, device=self.device)
            out = HeteroBatch(x, self.edge_index_dict(B), y, B)
            if s.extra_types:       # "r_o": [B, 4 * T] history of the (z-scored) body orientation; the batch carries its last row (quadSDKDataset_Morph.py:L346-348)
                out._window_extra = torch.empty(B, 4 * s.T, dtype=torch.float32, device=self.device)
        elif out.batch_size != B:
            raise ValueError("out has a different batch size")
        xs = out.x_dict
        bufs = [xs[t] for t in names] + ([out._window_extra] if s.extra_types else [])
        ptrs = (C.c_void_p * 4)(*([b.data_ptr() for b in bufs] + [None] * (4 - len(kt))))
        stream = torch.cuda.current_stream(self.device).cuda_stream
        with torch.cuda.device(self.device):
            N.check(N.lib().mshgnn_build_windows(C.byref(self._desc), self.seq.data_ptr(), self.labels.data_ptr(),
                                                 N.F64 if self.dtype == torch.float64 else N.F32, self.n_rows, idx.data_ptr(), B,
                                                 ptrs, out.y.data_ptr(), stream), "mshgnn_build_windows")
        if s.extra_types:
            out.r_o = out._window_extra.view(B, 4, s.T)[:, :, s.T - 1].reshape(-1)       # collated like PyG: [B * 4]
        return out
