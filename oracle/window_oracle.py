"""CPU restatement (numpy / torch fp64) of the reference's per-sample dataset path for the Mini Cheetah contact
dataset, used as the checker of the device-side window builder.

TEST INFRASTRUCTURE ONLY: imported by tests/, bench.py's checker leg and __graft_entry__.smoke(); the product path
(ms_hgnn.windows -> libmshgnn_b200.so) never imports it.

Follows, step by step and with the reference's own array operations:
  * ``LinTzuYaunDataset.load_data_at_dataset_seq``        src/ms_hgnn/datasets_py/LinTzuYaunDataset.py:L66-88
  * ``LinTzuYaunDataset.get_urdf_name_to_dataset_array_index``  LinTzuYaunDataset.py:L34-60   (+ URDF node order RL, FL, RR, FR,
    SURVEY Appendix A)  ->  ``joint_node_indices_sorted`` / ``foot_node_indices_sorted``  flexibleDataset.py:L131-142
  * ``LinTzuYaunDataset_Morph.load_data_sorted_k4 / _c2`` LinTzuYaunDataset_Morph.py:L156-347
  * ``FlexibleDataset.load_data_sorted``                  flexibleDataset.py:L336-400         (MI-HGNN: one base, no tiling)
  * ``LinTzuYaunDataset_Morph.apply_symmetry``            LinTzuYaunDataset_Morph.py:L349-408 (+ coefficient dicts L133-154)
  * ``get_helper_heterogeneous_gnn(_c2)``                 LinTzuYaunDataset_Morph.py:L555-697, flexibleDataset.py:L537-607
  * collate of x / y as torch_geometric ``Batch.from_data_list`` does (concatenate graph-major).

Pinning: the reference's only golden values for this path are the z-scored matrices of
``tests/testDatasets.py:L513-539`` (their raw inputs come from a dataset download that is not available offline).
``tests/test_windows.py`` checks that ``zscore`` leaves those matrices unchanged (mean 0 / Bessel std 1 to 1e-12), which
pins the ``correction=1`` choice; everything else here is PARITY UNPINNED (line-by-line restatement only).
"""
from __future__ import annotations

from typing import Dict, List, Optional

import numpy as np
import torch

# LinTzuYaunDataset.py:L41-59 with the URDF node order RL, FL, RR, FR
JOINT_NODE_INDICES_SORTED = np.array([9, 10, 11, 3, 4, 5, 6, 7, 8, 0, 1, 2], dtype=np.uint)
FOOT_NODE_INDICES_SORTED = np.array([3, 1, 2, 0], dtype=np.uint)
N_BASES = {"heterogeneous_gnn": 1, "heterogeneous_gnn_k4": 4, "heterogeneous_gnn_c2": 2}


def zscore(array: np.ndarray) -> np.ndarray:
    """flexibleDataset.py:L393-396 / LinTzuYaunDataset_Morph.py:L243-246: (x - mean) / std(correction=1), NaN -> 0 (and, as
    np.nan_to_num does, +-inf -> +-largest finite).  The reference divides a tensor by an ndarray inside np.nan_to_num(copy=False),
    which numpy >= 2 rejects; the quotient is formed in torch here, same fp64 arithmetic."""
    array_tensor = torch.from_numpy(np.ascontiguousarray(array, dtype=np.float64))
    q = (array_tensor - torch.mean(array_tensor, axis=0)) / torch.std(array_tensor, axis=0, correction=1)
    return np.nan_to_num(q.numpy(), copy=False, nan=0.0)


def _coefficients(group: dict, key: str, mode: str) -> Dict[str, np.ndarray]:
    refl = group[key]
    if mode == "MorphSym":          # create_morphsym_coefficients L140-154
        gs, gt = np.array(refl[0], dtype=np.float64), np.array(refl[1], dtype=np.float64)
        return {"gs": gs, "gt": gt, "gr": gs * gt}
    one = np.ones_like(refl[0], dtype=np.float64)      # create_coefficient_dict L133-138 ('Euclidean')
    return {"gs": one, "gt": one, "gr": one}


def _act(data: np.ndarray, perm, coef: Dict[str, np.ndarray], op: str) -> np.ndarray:
    """One branch of apply_symmetry (L361-406); works on [T, n] arrays and on the [4] label vector."""
    if op == "gs":
        return data[..., perm[0]].copy() * coef["gs"]
    if op == "gt":
        return data[..., perm[1]].copy() * coef["gt"]
    if op == "gr":
        data = data[..., perm[0]].copy()
        return data[..., perm[1]].copy() * coef["gr"]
    raise ValueError(op)


def sample(mat: Dict[str, np.ndarray], idx: int, model_type: str, T: int, normalize: bool = True,
           symmetry_operator: Optional[str] = None, symmetry_mode: Optional[str] = None, group: Optional[dict] = None):
    """One dataset entry -> (base_x, joint_x, foot_x, y) exactly as ``get_helper_heterogeneous_gnn*`` builds them."""
    nb = N_BASES[model_type]
    lin_acc = np.array(mat["imu_acc"][idx:idx + T]).reshape(T, 3)
    ang_vel = np.array(mat["imu_omega"][idx:idx + T]).reshape(T, 3)
    j_p = np.array(mat["q"][idx:idx + T]).reshape(T, 12)
    j_v = np.array(mat["qd"][idx:idx + T]).reshape(T, 12)
    f_p = np.array(mat["p"][idx:idx + T]).reshape(T, 12)
    f_v = np.array(mat["v"][idx:idx + T]).reshape(T, 12)
    labels = np.squeeze(np.array(mat["contacts"][idx:idx + T])[-1])
    if nb > 1:
        lin_acc, ang_vel = np.tile(lin_acc, (1, nb)), np.tile(ang_vel, (1, nb))
    base = [lin_acc, ang_vel]
    sym = symmetry_operator is not None
    if sym:
        if nb == 1:
            raise ValueError("the reference applies dataset symmetries only in the _Morph datasets")
        base = [_act(base[0], group["permutation_Q_bs"], _coefficients(group, "reflection_Q_bs_lin", symmetry_mode), symmetry_operator),
                _act(base[1], group["permutation_Q_bs"], _coefficients(group, "reflection_Q_bs_ang", symmetry_mode), symmetry_operator)]
    joints = [a[:, JOINT_NODE_INDICES_SORTED] for a in (j_p, j_v)]
    if sym:
        cj = _coefficients(group, "reflection_Q_js", symmetry_mode)
        joints = [_act(a, group["permutation_Q_js"], cj, symmetry_operator) for a in joints]
    sorted_indices = [int(index * 3 + i) for index in FOOT_NODE_INDICES_SORTED for i in range(3)]
    feet = [a[:, sorted_indices] for a in (f_p, f_v)]
    if sym:
        cf = _coefficients(group, "reflection_Q_fs", symmetry_mode)
        feet = [_act(a, group["permutation_Q_fs"], cf, symmetry_operator) for a in feet]
    labels_sorted = labels[FOOT_NODE_INDICES_SORTED]
    if sym:
        labels_sorted = _act(labels_sorted, group["permutation_Q_ls"], _coefficients(group, "reflection_Q_ls", symmetry_mode), symmetry_operator)
    if normalize:
        if T <= 1:
            raise ValueError("normalize needs history_length > 1 (the reference returns None arrays)")
        base, joints, feet = [zscore(a) for a in base], [zscore(a) for a in joints], [zscore(a) for a in feet]
    base_x = torch.ones((nb, 2 * 3 * T), dtype=torch.float64)
    joint_x = torch.ones((12, 2 * T), dtype=torch.float64)
    foot_x = torch.ones((4, 2 * 3 * T), dtype=torch.float64)
    for i in range(nb):
        base_x[i] = torch.cat([torch.tensor(np.asarray(base[k])[:, i * 3:(i + 1) * 3].flatten("F"), dtype=torch.float64) for k in range(2)])
    for i in range(12):
        joint_x[i] = torch.cat([torch.tensor(np.asarray(joints[k])[:, i].flatten("F"), dtype=torch.float64) for k in range(2)])
    for i in range(4):
        foot_x[i] = torch.cat([torch.tensor(np.asarray(feet[k])[:, 3 * i:3 * i + 3].flatten("F"), dtype=torch.float64) for k in range(2)])
    return base_x, joint_x, foot_x, torch.tensor(np.asarray(labels_sorted, dtype=np.float64), dtype=torch.float64)


def batch(mat: Dict[str, np.ndarray], indices: List[int], model_type: str, T: int, **kw):
    """Collated node features and labels of the entries ``indices`` (graph-major concatenation)."""
    parts = [sample(mat, int(i), model_type, T, **kw) for i in indices]
    return {"base": torch.cat([p[0] for p in parts]), "joint": torch.cat([p[1] for p in parts]),
            "foot": torch.cat([p[2] for p in parts])}, torch.cat([p[3] for p in parts])


def synthetic_mat(n_rows: int, seed: int = 0, dtype=np.float64) -> Dict[str, np.ndarray]:
    """A random stand-in for the dataset's data.mat (same keys and shapes, LinTzuYaunDataset.py:L79-86): offset random
    walks for the kinematic channels, one constant column (exercises the NaN -> 0 branch), Bernoulli contacts."""
    rng = np.random.default_rng(seed)
    def walk(c, scale, offset):
        return (offset + np.cumsum(rng.normal(0.0, scale, size=(n_rows, c)), axis=0) * 0.05 + rng.normal(0.0, scale, size=(n_rows, c))).astype(dtype)
    mat = {"imu_acc": walk(3, 1.0, np.array([0.0, 0.0, 9.8])), "imu_omega": walk(3, 0.3, 0.0), "q": walk(12, 0.2, 0.6),
           "qd": walk(12, 2.0, 0.0), "p": walk(12, 0.05, -0.25), "v": walk(12, 0.5, 0.0),
           "tau_est": walk(12, 1.0, 0.0), "contacts": (rng.random((n_rows, 4)) < 0.5).astype(dtype)}
    mat["p"][:, 7] = dtype(-0.125)
    return mat
