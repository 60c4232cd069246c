"""CPU restatement (numpy / torch fp64) of the reference's per-sample dataset paths - Mini Cheetah contact data
(``sample``), quad-SDK A1 ground-reaction-force data (``sample_a1``), Solo12 centre-of-mass data (``sample_solo``) - used as
the checker of the device-side window builder.

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


def sample_a1(mat: Dict[str, np.ndarray], idx: int, T: int, normalize: bool = True, grf_dimension: int = 1,
              grf_body_to_world_frame: bool = False, symmetry_operator: Optional[str] = None, symmetry_mode: Optional[str] = None,
              group: Optional[dict] = None):
    """One entry of QuadSDKDataset_A1 with model_type 'heterogeneous_gnn_c2' -> (base_x, joint_x, foot_x, y, r_o or None).
    Follows quadSDKDataset_Morph.py: load_data_at_dataset_seq_3d / load_data_at_dataset_seq L444-489 (label rotation by the last
    orientation with scipy), load_data_sorted_c2 L100-174 (base tiling, identity URDF order L404-441, apply_symmetry L177-239, z-score
    of base / joint arrays and of r_o), get_helper_heterogeneous_gnn_c2 L274-351 (widths flexibleDataset.py:L185-190: foot_x stays ones)."""
    from scipy.spatial.transform import Rotation
    lin_acc = np.array(mat["imu_acc"][idx:idx + T]).reshape(T, 3)
    ang_vel = np.array(mat["imu_omega"][idx:idx + T]).reshape(T, 3)
    j_p = np.array(mat["q"][idx:idx + T]).reshape(T, 12)
    j_v = np.array(mat["qd"][idx:idx + T]).reshape(T, 12)
    j_T = np.array(mat["tau"][idx:idx + T]).reshape(T, 12)
    grfs = np.squeeze(np.array(mat["F"][idx:idx + T]).reshape(T, 12))[-1] if T > 1 else np.array(mat["F"][idx:idx + T]).reshape(12)
    r_quat = np.array(mat["r_o"][idx:idx + T]).reshape(T, 4)
    if grf_body_to_world_frame:
        world_to_body_R = Rotation.from_quat(r_quat[-1])
        grfs_T = np.array(grfs.reshape(4, 3), dtype=np.double).T
        grfs = (world_to_body_R.as_matrix() @ grfs_T).T.flatten()
    labels = grfs[[2, 5, 8, 11]] if grf_dimension == 1 else grfs
    base = [np.tile(lin_acc, (1, 2)), np.tile(ang_vel, (1, 2))]
    sym = symmetry_operator is not None
    if sym:
        base = [_act(base[0], group["permutation_Q_bs"], _coefficients(group, "reflection_Q_bs_lin", symmetry_mode), symmetry_operator),
                _act(base[1], group["permutation_Q_bs"], _coefficients(group, "reflection_Q_bs_ang", symmetry_mode), symmetry_operator)]
    order = np.arange(12, dtype=np.uint)                 # joint_node_indices_sorted for a1_pruned.urdf (identity)
    joints = [a[:, order] for a in (j_p, j_v, j_T)]
    if sym:
        cj = _coefficients(group, "reflection_Q_js", symmetry_mode)
        joints = [_act(a, group["permutation_Q_js"], cj, symmetry_operator) for a in joints]
    foot_order = np.arange(4, dtype=np.uint)
    if grf_dimension == 1:
        labels_sorted = labels[foot_order]
        if sym:
            labels_sorted = _act(labels_sorted, group["permutation_Q_ls"], _coefficients(group, "reflection_Q_ls", symmetry_mode), symmetry_operator)
    else:
        labels_sorted = labels[[int(index * 3 + i) for index in foot_order for i in range(3)]]
        if sym:
            labels_sorted = _act(labels_sorted, group["permutation_Q_fs"], _coefficients(group, "reflection_Q_fs", symmetry_mode), symmetry_operator)
    r_o = r_quat
    if normalize:
        if T <= 1:
            raise ValueError("normalize needs history_length > 1 (the reference returns None arrays)")
        base, joints, r_o = [zscore(a) for a in base], [zscore(a) for a in joints], zscore(r_quat)
    base_x = torch.ones((2, 2 * 3 * T), dtype=torch.float64)
    joint_x = torch.ones((12, 3 * T), dtype=torch.float64)
    foot_x = torch.ones((4, 1), dtype=torch.float64)
    for i in range(2):
        base_x[i] = torch.cat([torch.tensor(np.asarray(base[k])[:, i * 3:(i + 1) * 3].flatten("F"), dtype=torch.float64) for k in range(2)])
    for i in range(12):
        joint_x[i] = torch.cat([torch.tensor(np.asarray(joints[k])[:, i].flatten("F"), dtype=torch.float64) for k in range(3)])
    y = torch.tensor(np.asarray(labels_sorted, dtype=np.float64), dtype=torch.float64)
    return base_x, joint_x, foot_x, y, (torch.tensor(np.asarray(r_o)[-1], dtype=torch.float64) if grf_body_to_world_frame else None)


def sample_solo(mat: Dict[str, np.ndarray], idx: int, model_type: str, T: int, normalize: bool = True, joint_order=None):
    """One entry of Solo12Dataset -> (base_x, joint_x, y).  Follows soloDataset.py: dataset-level standardisation L136-143 with the
    stored statistics (Standarizer.transform L18-31), load_data_at_dataset_seq L382-401 (zero base inputs, label = last frame of Y),
    load_data_sorted / _k4 / _c2 L332-380, L546-718 (base tiling, joint order, labels repeated per base node and interleaved
    [lin | ang]), get_helper_heterogeneous_gnn L235-300.  joint_order = joint_node_indices_sorted (the Solo URDF is not in the tree)."""
    nb = {"heterogeneous_gnn_k4_com": 4, "heterogeneous_gnn_c2_com": 2, "heterogeneous_gnn_s4_com": 1}[model_type]
    X, Y = np.asarray(mat["X"], dtype=np.float64), np.asarray(mat["Y"], dtype=np.float64)
    if normalize:
        X, Y = (X - mat["x_mean"]) / mat["x_std"], (Y - mat["y_mean"]) / mat["y_std"]
    j_p = X[:, :12][idx:idx + T].reshape(T, 12)
    j_v = X[:, 12:][idx:idx + T].reshape(T, 12)
    base_lin_vel = Y[:, :3][idx:idx + T].reshape(T, 3)
    base_ang_vel = Y[:, 3:][idx:idx + T].reshape(T, 3)
    labels = np.concatenate([base_lin_vel[-1], base_ang_vel[-1]])
    lin_vel, ang_vel = np.zeros((T, 3)), np.zeros((T, 3))
    if nb > 1:
        lin_vel, ang_vel = np.tile(lin_vel, (1, nb)), np.tile(ang_vel, (1, nb))
    order = np.arange(12) if joint_order is None else np.asarray(joint_order)
    joints = [a[:, order] for a in (j_p, j_v)]
    if nb == 1:
        labels_sorted = labels
    else:
        lin, ang = np.tile(labels[:3], nb), np.tile(labels[3:], nb)
        labels_sorted = [part[3 * i:3 * i + 3] for i in range(nb) for part in (lin, ang)]
    base_x = torch.ones((nb, 6 * T), dtype=torch.float64)
    joint_x = torch.ones((12, 2 * T), dtype=torch.float64)
    base = [lin_vel, ang_vel]
    for i in range(nb):
        base_x[i] = torch.cat([torch.tensor(base[k][:, i * 3:(i + 1) * 3].flatten("F"), dtype=torch.float64) for k in range(2)])
    for i in range(12):
        joint_x[i] = torch.cat([torch.tensor(joints[k][:, i].flatten("F"), dtype=torch.float64) for k in range(2)])
    return base_x, joint_x, torch.tensor(np.array(labels_sorted).reshape(-1), dtype=torch.float64)


def synthetic_a1_mat(n_rows: int, seed: int = 0, dtype=np.float64) -> Dict[str, np.ndarray]:
    """Random stand-in for the quad-SDK data.mat (keys and shapes of quadSDKDataset.py:L107-117); unit-ish quaternions."""
    rng = np.random.default_rng(seed)
    def walk(c, scale, offset):
        return (offset + np.cumsum(rng.normal(0.0, scale, size=(n_rows, c)), axis=0) * 0.05 + rng.normal(0.0, scale, size=(n_rows, c))).astype(dtype)
    q = walk(4, 0.05, np.array([0.02, -0.03, 0.1, 0.99]))
    return {"imu_acc": walk(3, 1.0, np.array([0.0, 0.0, 9.8])), "imu_omega": walk(3, 0.3, 0.0), "q": walk(12, 0.2, 0.6), "qd": walk(12, 2.0, 0.0),
            "tau": walk(12, 3.0, 0.0), "F": walk(12, 10.0, 30.0), "r_p": walk(3, 0.1, 0.3), "r_o": q, "timestamps": np.zeros((n_rows, 3), dtype=dtype)}


def synthetic_solo_mat(n_rows: int, seed: int = 0) -> Dict[str, np.ndarray]:
    rng = np.random.default_rng(seed)
    X, Y = rng.normal(0.3, 1.5, size=(n_rows, 24)), rng.normal(-0.1, 0.7, size=(n_rows, 6))
    return {"X": X, "Y": Y, "x_mean": X.mean(0), "x_std": X.std(0), "y_mean": Y.mean(0), "y_std": Y.std(0)}


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
