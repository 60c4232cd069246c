"""Morphology graph templates and symmetry-group tables.

The reference builds these on the host in its dataset classes; here they are constants, because
the per-graph ``edge_index`` of every batch is fully determined by (template, batch size):

* base kinematic edges: ``graphParser.py:L483-550`` (pinned by ``tests/testGraphParser.py:L370-374``)
* K4 Mini Cheetah: ``datasets_py/LinTzuYaunDataset_Morph.py:L410-444`` + metadata ``L531-540``
* C2 Mini Cheetah: ``LinTzuYaunDataset_Morph.py:L492-523`` + ``L541-551``
* C2 A1: ``datasets_py/quadSDKDataset_Morph.py:L241-272``
* K4 / C2 / S4 Solo (COM): ``datasets_py/soloDataset.py:L201-233, L455-521``
* MI-HGNN baseline: ``datasets_py/flexibleDataset.py:L317-323``
Group tables: ``cfg/*.yaml`` of the reference (values restated in ``GROUPS`` below and written out
as yaml by ``tools/make_cfg.py``).
"""
from __future__ import annotations

import os
from dataclasses import dataclass
from typing import Dict, List, Sequence, Tuple

import torch

EdgeType = Tuple[str, str, str]

# kinematic chain of a 4-legged, 3-joints-per-leg robot (joint 3k hip, 3k+1 thigh, 3k+2 calf)
_JJ = ([0, 1, 1, 2, 3, 4, 4, 5, 6, 7, 7, 8, 9, 10, 10, 11],
       [1, 0, 2, 1, 4, 3, 5, 4, 7, 6, 8, 7, 10, 9, 11, 10])
_FJ = ([0, 1, 2, 3], [2, 5, 8, 11])
_JF = (_FJ[1], _FJ[0])


def _rev(e):
    return (list(e[1]), list(e[0]))


@dataclass(frozen=True)
class Template:
    name: str
    node_types: Tuple[str, ...]
    nodes_per_graph: Dict[str, int]
    edge_types: Tuple[EdgeType, ...]
    edges: Dict[EdgeType, Tuple[List[int], List[int]]]   # per-graph (src list, dst list)

    @property
    def metadata(self):
        return list(self.node_types), [tuple(e) for e in self.edge_types]

    def edge_index(self, et: EdgeType, B: int, device=None) -> torch.Tensor:
        """Batched edge_index [2, E*B] in PyG ``Batch.from_data_list`` layout."""
        src, dst = self.edges[et]
        t = torch.tensor([src, dst], dtype=torch.long)
        E = t.shape[1]
        g = torch.arange(B, dtype=torch.long).repeat_interleave(E)
        t = t.repeat(1, B)
        out = torch.stack((t[0] + g * self.nodes_per_graph[et[0]], t[1] + g * self.nodes_per_graph[et[2]]))
        return out.to(device) if device is not None else out

    def edge_index_dict(self, B: int, device=None) -> Dict[EdgeType, torch.Tensor]:
        return {et: self.edge_index(et, B, device) for et in self.edge_types}


def _mk(name, node_types, counts, edge_list) -> Template:
    return Template(name, tuple(node_types), dict(zip(node_types, counts)),
                    tuple(e for e, _ in edge_list), {e: (list(v[0]), list(v[1])) for e, v in edge_list})


_BJ_K4 = ([0, 1, 2, 3], [0, 3, 6, 9])

K4_MINI_CHEETAH = _mk("k4_mini_cheetah", ("base", "joint", "foot"), (4, 12, 4), [
    (("base", "connect", "joint"), _BJ_K4),
    (("joint", "connect", "base"), _rev(_BJ_K4)),
    (("joint", "connect", "joint"), _JJ),
    (("foot", "connect", "joint"), _FJ),
    (("joint", "connect", "foot"), _JF),
    (("base", "gt", "base"), ([0, 1, 2, 3], [1, 0, 3, 2])),
    (("base", "gs", "base"), ([0, 2, 1, 3], [2, 0, 3, 1])),
])

C2_MINI_CHEETAH = _mk("c2_mini_cheetah", ("base", "joint", "foot"), (2, 12, 4), [
    (("base", "front_bj", "joint"), ([0, 1], [3, 9])),
    (("joint", "front_bj", "base"), ([3, 9], [0, 1])),
    (("base", "back_bj", "joint"), ([0, 1], [0, 6])),
    (("joint", "back_bj", "base"), ([0, 6], [0, 1])),
    (("joint", "connect", "joint"), _JJ),
    (("foot", "connect", "joint"), _FJ),
    (("joint", "connect", "foot"), _JF),
    (("base", "center_bb", "base"), ([0, 1], [1, 0])),
])

C2_A1 = _mk("c2_a1", ("base", "joint", "foot"), (2, 12, 4), [
    (("base", "front_bj", "joint"), ([0, 1], [0, 6])),
    (("joint", "front_bj", "base"), ([0, 6], [0, 1])),
    (("base", "back_bj", "joint"), ([0, 1], [3, 9])),
    (("joint", "back_bj", "base"), ([3, 9], [0, 1])),
    (("joint", "connect", "joint"), _JJ),
    (("foot", "connect", "joint"), _FJ),
    (("joint", "connect", "foot"), _JF),
    (("base", "center_bb", "base"), ([0, 1], [1, 0])),
])

# Solo: note gs / gt edge lists are swapped w.r.t. Mini Cheetah (soloDataset.py:L476-479)
K4_SOLO_COM = _mk("k4_solo_com", ("base", "joint"), (4, 12), [
    (("base", "connect", "joint"), _BJ_K4),
    (("joint", "connect", "base"), _rev(_BJ_K4)),
    (("joint", "connect", "joint"), _JJ),
    (("base", "gt", "base"), ([0, 2, 1, 3], [2, 0, 3, 1])),
    (("base", "gs", "base"), ([0, 1, 2, 3], [1, 0, 3, 2])),
])

C2_SOLO_COM = _mk("c2_solo_com", ("base", "joint"), (2, 12), [
    (("base", "front_bj", "joint"), ([0, 1], [3, 9])),
    (("joint", "front_bj", "base"), ([3, 9], [0, 1])),
    (("base", "back_bj", "joint"), ([0, 1], [0, 6])),
    (("joint", "back_bj", "base"), ([0, 6], [0, 1])),
    (("joint", "connect", "joint"), _JJ),
    (("base", "center_bb", "base"), ([0, 1], [1, 0])),
])

_BJ_MI = ([0, 0, 0, 0], [0, 3, 6, 9])
MI_QUADRUPED = _mk("mi_quadruped", ("base", "joint", "foot"), (1, 12, 4), [
    (("base", "connect", "joint"), _BJ_MI),
    (("joint", "connect", "base"), _rev(_BJ_MI)),
    (("joint", "connect", "joint"), _JJ),
    (("foot", "connect", "joint"), _FJ),
    (("joint", "connect", "foot"), _JF),
])

S4_SOLO_COM = _mk("s4_solo_com", ("base", "joint"), (1, 12), [
    (("base", "connect", "joint"), _BJ_MI),
    (("joint", "connect", "base"), _rev(_BJ_MI)),
    (("joint", "connect", "joint"), _JJ),
])

TEMPLATES = {t.name: t for t in (K4_MINI_CHEETAH, C2_MINI_CHEETAH, C2_A1, K4_SOLO_COM, C2_SOLO_COM, MI_QUADRUPED, S4_SOLO_COM)}


# ------------------------------------------------------------------------------------------
# group tables.  Rows are [gs (sagittal), gt (transversal)]; per-leg reflection triples tiled x4.
# ------------------------------------------------------------------------------------------
def _tile(v, n):
    return [int(x) for x in list(v) * n]


_PERM12 = [[6, 7, 8, 9, 10, 11, 0, 1, 2, 3, 4, 5], [3, 4, 5, 0, 1, 2, 9, 10, 11, 6, 7, 8]]
_JS = [_tile([-1, 1, 1], 4), _tile([1, -1, -1], 4)]
_LIN = lambda n: [_tile([1, -1, 1], n), _tile([-1, 1, 1], n)]
_ANG = lambda n: [_tile([-1, 1, -1], n), _tile([1, -1, -1], n)]
_LS = {"permutation_Q_ls": [[2, 3, 0, 1], [1, 0, 3, 2]], "reflection_Q_ls": [[1, 1, 1, 1], [1, 1, 1, 1]]}


def _k4_quadruped():
    return {"group_label": "K4", "permutation_Q_js": _PERM12, "reflection_Q_js": _JS,
            "permutation_Q_bs": _PERM12, "reflection_Q_bs_lin": _LIN(4), "reflection_Q_bs_ang": _ANG(4),
            "permutation_Q_fs": _PERM12, "reflection_Q_fs": _LIN(4), **_LS}


def _c2_quadruped():
    return {"group_label": "C2", "permutation_Q_js": _PERM12, "reflection_Q_js": _JS,
            "permutation_Q_bs": [[3, 4, 5, 0, 1, 2], [0, 1, 2, 3, 4, 5]],
            "reflection_Q_bs_lin": _LIN(2), "reflection_Q_bs_ang": _ANG(2),
            "permutation_Q_fs": _PERM12, "reflection_Q_fs": _LIN(4), **_LS}


def _k4_solo():
    # cfg/solo-k4.yaml / solo12-k4.yaml: the foot-space block is commented out in the reference
    return {"group_label": "K4", "permutation_Q_js": _PERM12, "reflection_Q_js": _JS,
            "permutation_Q_bs": _PERM12, "reflection_Q_bs_lin": _LIN(4), "reflection_Q_bs_ang": _ANG(4),
            "permutation_Q_ls": _PERM12, "reflection_Q_ls_lin": _LIN(4), "reflection_Q_ls_ang": _ANG(4)}


GROUPS = {
    "mini_cheetah-k4": _k4_quadruped(),
    "mini_cheetah-c2": _c2_quadruped(),
    "a1-c2": _c2_quadruped(),
    "solo-c2": _c2_quadruped(),
    "solo-k4": _k4_solo(),
    "solo12-k4": _k4_solo(),
}

CFG_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "cfg")


def cfg_path(name: str) -> str:
    """Path of the packaged yaml group table, e.g. cfg_path('mini_cheetah-k4')."""
    p = os.path.join(CFG_DIR, name + ".yaml")
    if not os.path.exists(p):
        raise FileNotFoundError(p)
    return p


def load_group(path: str) -> dict:
    import yaml
    with open(path, "r") as f:
        return yaml.safe_load(f)


# ------------------------------------------------------------------------------------------
# +-1 tables of the models (hgnn_k4.py:L37-94, hgnn_c2.py:L42-83, *_com.py)
# ------------------------------------------------------------------------------------------
def k4_sign_tables(group, with_feet=True) -> Dict[str, torch.Tensor]:
    """Per local leg/base index k = 0..3 -> [e, gt, gs, gs*gt]."""
    one = torch.ones(3, dtype=torch.float64)

    def quad(key):
        if group is None:
            return torch.ones(12, dtype=torch.float64)
        rows = group[key]          # KeyError/TypeError when the block is missing, as in the reference
        gs = torch.tensor(rows[0][:3], dtype=torch.float64)
        gt = torch.tensor(rows[1][:3], dtype=torch.float64)
        return torch.cat((one, gt, gs, gs * gt))

    out = {"joint": quad("reflection_Q_js")}
    if with_feet:
        out["foot"] = quad("reflection_Q_fs")
    out["base_lin"] = quad("reflection_Q_bs_lin")
    out["base_ang"] = quad("reflection_Q_bs_ang")
    return out


def c2_sign_tables(group, with_feet=True) -> Dict[str, torch.Tensor]:
    """Legs [e, e, gs, gs]; bases [e, gs]."""
    one = torch.ones(3, dtype=torch.float64)

    def gs(key):
        if group is None:
            return one.clone()
        return torch.tensor(group[key][0][:3], dtype=torch.float64)

    js = gs("reflection_Q_js")
    out = {"joint": torch.cat((one, one, js, js))}
    if with_feet:
        fs = gs("reflection_Q_fs")
        out["foot"] = torch.cat((one, one, fs, fs))
    out["base_lin"] = torch.cat((one, gs("reflection_Q_bs_lin")))
    out["base_ang"] = torch.cat((one, gs("reflection_Q_bs_ang")))
    return out


def blockwise_signs(n_nodes: int, first: torch.Tensor, second: torch.Tensor, T: int) -> List[float]:
    """Row layout [2 variables][3 axes][T steps]: first variable * first[3n+d], second * second[3n+d]."""
    out: List[float] = []
    for n in range(n_nodes):
        for tab in (first, second):
            for d in range(3):
                out += [float(tab[3 * n + d])] * T
    return out


def rowwise_signs(per_node: Sequence[float], width: int) -> List[float]:
    out: List[float] = []
    for s in per_node:
        out += [float(s)] * width
    return out
