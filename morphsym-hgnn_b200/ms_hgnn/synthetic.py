"""Synthetic batches with the exact layout of the reference's collated ``HeteroDataBatch``.

Layout: node tensors concatenated graph-major, ``edge_index`` per edge type tiled
with per-graph offsets (PyG ``Batch.from_data_list``), ``y`` concatenated to ``[B*|y|]``,
``batch_size == B``.  Used by the tests, ``bench.py`` and ``__graft_entry__.smoke()``; there are no
datasets in this image, so all measured inputs are synthetic windows of each config's shape.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, Optional, Tuple

import torch

from . import morphology as M

EdgeType = Tuple[str, str, str]


class HeteroBatch:
    """Duck-typed stand-in for torch_geometric's HeteroDataBatch (only what the hot path reads)."""

    def __init__(self, x: Dict[str, torch.Tensor], edge_index: Dict[EdgeType, torch.Tensor], y: torch.Tensor,
                 batch_size: int, r_o: Optional[torch.Tensor] = None):
        self._x = x
        self._edge_index = edge_index
        self.y = y
        self.batch_size = batch_size
        if r_o is not None:
            self.r_o = r_o

    @property
    def x_dict(self) -> Dict[str, torch.Tensor]:
        return dict(self._x)          # a fresh dict per access, like PyG

    @property
    def edge_index_dict(self) -> Dict[EdgeType, torch.Tensor]:
        return dict(self._edge_index)

    def to(self, device, non_blocking: bool = False):
        return HeteroBatch({k: v.to(device, non_blocking=non_blocking) for k, v in self._x.items()},
                           {k: v.to(device, non_blocking=non_blocking) for k, v in self._edge_index.items()},
                           self.y.to(device, non_blocking=non_blocking), self.batch_size,
                           getattr(self, "r_o", None).to(device) if hasattr(self, "r_o") else None)

    def shard(self, rank: int, world: int) -> "HeteroBatch":
        """Contiguous graph shard for data parallelism: every rank's node rows are a contiguous
        slice of each x tensor and its edge_index is the same template tiled over the smaller batch."""
        B = self.batch_size
        per = (B + world - 1) // world
        g0, g1 = min(rank * per, B), min((rank + 1) * per, B)
        nb = g1 - g0
        if nb <= 0:
            raise ValueError(f"rank {rank} of {world} gets no graphs out of {B}")
        x = {}
        for k, v in self._x.items():
            n = v.shape[0] // B
            x[k] = v[g0 * n:g1 * n]
        ei = {}
        for k, v in self._edge_index.items():
            E = v.shape[1] // B
            ei[k] = v[:, :E * nb]          # the first nb graphs' block IS the template tiled nb times from offset 0
        w = self.y.numel() // B
        return HeteroBatch(x, ei, self.y.reshape(-1)[g0 * w:g1 * w], nb)

    def pin_memory(self):
        return HeteroBatch({k: v.pin_memory() for k, v in self._x.items()},
                           {k: v.pin_memory() for k, v in self._edge_index.items()},
                           self.y.pin_memory(), self.batch_size)


@dataclass(frozen=True)
class Config:
    """One BASELINE.json workload."""
    name: str
    model: str                 # class name in ms_hgnn
    template: str              # key of morphology.TEMPLATES
    group: Optional[str]       # packaged yaml name
    in_width: Dict[str, int]
    regression: bool
    grf_dimension: int = 1
    loss: str = "mse"          # "mse" | "ce"
    label_width: int = 4       # per graph
    zscore: bool = True


CONFIGS = {
    # BASELINE.json configs[0..4]
    "a1-c2-grf": Config("a1-c2-grf", "GRF_HGNN_C2", "c2_a1", "a1-c2", {"base": 900, "joint": 450, "foot": 1}, True, 3, "mse", 12, False),
    "mini_cheetah-c2-contact": Config("mini_cheetah-c2-contact", "GRF_HGNN_C2", "c2_mini_cheetah", "mini_cheetah-c2",
                                      {"base": 900, "joint": 300, "foot": 900}, False, 1, "ce", 4),
    "mini_cheetah-k4-contact": Config("mini_cheetah-k4-contact", "GRF_HGNN_K4", "k4_mini_cheetah", "mini_cheetah-k4",
                                      {"base": 900, "joint": 300, "foot": 900}, False, 1, "ce", 4),
    # config 4 does not exist in the reference (solo-k4.yaml lacks reflection_Q_fs); defined as the K4 GRF
    # regression head on the K4 quadruped template with the Mini Cheetah table (identical numbers), SURVEY 8d.
    "k4-grf-regression": Config("k4-grf-regression", "GRF_HGNN_K4", "k4_mini_cheetah", "mini_cheetah-k4",
                                {"base": 900, "joint": 300, "foot": 900}, True, 1, "mse", 4, False),
    "solo12-k4-com": Config("solo12-k4-com", "COM_HGNN_K4", "k4_solo_com", "solo12-k4", {"base": 6, "joint": 2}, True, 1, "mse", 24, False),
    # extra parity cases
    "solo-c2-com": Config("solo-c2-com", "COM_HGNN_C2", "c2_solo_com", "solo-c2", {"base": 6, "joint": 2}, True, 1, "mse", 12, False),
    "mi-grf": Config("mi-grf", "GRF_HGNN", "mi_quadruped", None, {"base": 6, "joint": 3, "foot": 1}, True, 1, "mse", 4, False),
    "mi-contact": Config("mi-contact", "GRF_HGNN", "mi_quadruped", None, {"base": 900, "joint": 300, "foot": 900}, False, 1, "ce", 4),
    "mi-com": Config("mi-com", "COM_HGNN", "s4_solo_com", None, {"base": 6, "joint": 2}, True, 1, "mse", 6, False),
}


def _windows(n_rows: int, width: int, zscore: bool, gen: torch.Generator, dtype) -> torch.Tensor:
    x = torch.randn(n_rows, width, generator=gen, dtype=torch.float64)
    if zscore and width % 150 == 0:
        v = x.view(n_rows, width // 150, 150)
        v = (v - v.mean(-1, keepdim=True)) / v.std(-1, keepdim=True)       # Bessel, flexibleDataset.py:L390-398
        x = v.reshape(n_rows, width)
    return x.to(dtype)


def make_batch(cfg: Config, B: int, seed: int = 0, dtype=torch.float32, device="cpu") -> HeteroBatch:
    """Seeded synthetic batch of B windows in the reference's collated layout (created on the host)."""
    tpl = M.TEMPLATES[cfg.template]
    gen = torch.Generator().manual_seed(1000 + seed)
    x = {}
    for t in tpl.node_types:
        n = tpl.nodes_per_graph[t]
        w = cfg.in_width[t]
        if t == "base" and w == 900:
            # the same IMU window tiled over the base nodes (LinTzuYaunDataset_Morph.py:L192-193)
            one = _windows(B, w, cfg.zscore, gen, dtype)
            x[t] = one.repeat_interleave(n, dim=0)
        elif t == "base" and cfg.model.startswith("COM"):
            x[t] = torch.zeros(B * n, w, dtype=dtype)                      # soloDataset.py:L396-397
        elif t == "foot" and w == 1:
            x[t] = torch.ones(B * n, 1, dtype=dtype)                       # A1: constant foot feature
        else:
            x[t] = _windows(B * n, w, cfg.zscore, gen, dtype)
    if cfg.loss == "ce":
        y = (torch.rand(B * cfg.label_width, generator=gen) > 0.5).to(dtype)
    elif cfg.model.startswith("COM"):
        n_base = tpl.nodes_per_graph["base"]
        one = torch.randn(B, 6, generator=gen, dtype=torch.float64).to(dtype)
        y = one.repeat(1, n_base).reshape(-1)                               # 6-vector tiled over bases
    else:
        y = (torch.rand(B * cfg.label_width, generator=gen, dtype=torch.float64) * 70.0).to(dtype)
    batch = HeteroBatch(x, tpl.edge_index_dict(B), y, B)
    return batch.to(device) if str(device) != "cpu" else batch


def build_model(cfg: Config, hidden: int = 128, layers: int = 8, seed: int = 0, module=None):
    """Instantiate cfg.model from ``module`` (default: the native ms_hgnn package) with seeded weights."""
    if module is None:
        import ms_hgnn as module
    tpl = M.TEMPLATES[cfg.template]
    cls = getattr(module, cfg.model)
    kw = dict(hidden_channels=hidden, num_layers=layers, data_metadata=tpl.metadata, regression=cfg.regression,
              in_dims=dict(cfg.in_width))
    if cfg.group is not None:
        kw.update(symmetry_mode="MorphSym", group_operator_path=M.cfg_path(cfg.group))
    if cfg.model in ("GRF_HGNN_C2", "GRF_HGNN"):
        kw["grf_dimension"] = cfg.grf_dimension
    torch.manual_seed(2000 + seed)
    return cls(**kw)
