"""CPU oracle for the MS-HGNN forward/backward hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is part of the product: only
``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference``
legs of ``bench.py`` may import it, and there only as the checker / CPU timing arm.

What it is
----------
A plain-PyTorch (CPU, float64 by default) restatement of the op sequence the reference
executes through ``torch_geometric==2.5.0`` (pinned in the reference's
``pyproject.toml:L16``; not vendored, not installable offline).  The arithmetic of
``HeteroDictLinear`` / ``HeteroConv`` / ``GraphConv`` is restated from PyG's published
semantics:

* ``GraphConv``:   out_i = lin_rel( AGG_{j->i} x_j ) + lin_root( x_i ); ``lin_rel`` has a bias,
  ``lin_root`` has none; AGG in {add, mean}; mean = sum / clamp(in_degree, 1).
* ``HeteroConv(aggr='sum')``: per destination type, sum of the per-edge-type outputs in
  metadata order.
* ``HeteroDictLinear``: one independent ``Linear`` (weight [H, in], bias [H]) per node type.
* PyG ``Linear.reset_parameters``: kaiming_uniform(a=sqrt(5)) == U(+-1/sqrt(fan_in)) for the
  weight and U(+-1/sqrt(fan_in)) for the bias.

Reference call sites this follows (paths relative to /root/reference/src/ms_hgnn/lightning_py):
  hgnn.py:L5-63 (GRF_HGNN), hgnn.py:L66-118 (COM_HGNN), hgnn_k4.py:L10-196 (GRF_HGNN_K4),
  hgnn_k4.py:L198-289 (apply_symmetry), hgnn_c2.py:L10-189 (GRF_HGNN_C2, ms_foot_decoder),
  hgnn_c2.py:L191-284, hgnn_k4_com.py:L10-177 (COM_HGNN_K4), hgnn_c2_com.py:L10-172,
  hgnn_s4_com.py:L6-71, gnnLightning.py:L124-151 / L633-660 (loss heads),
  customMetrics.py:L6-25 (CE = sum / N, rounded through float32), gnnLightning_com.py:L96-121.

Pinning
-------
PINNED against the reference's own code, three ways:

1. ``tests/test_reference_pin.py``: the reference's UNMODIFIED model files (``hgnn.py``, ``hgnn_k4.py``, ``hgnn_c2.py``,
   ``hgnn_k4_com.py``, ``hgnn_c2_com.py``, ``hgnn_s4_com.py``) are imported from /root/reference against
   ``oracle/pyg_shim`` (the four ``torch_geometric.nn`` classes they use, restated from torch_geometric 2.5.0) and run
   in fp64; every class of this file equals them on every ``CONFIGS`` entry at L = 8 - outputs, loss and EVERY
   gradient tensor - to fp64 rounding (<= 1e-12 norm-wise; summation order only), with identical
   ``named_parameters()`` order and identical dead-gradient sets.  That covers the sign tables, ``base_transform`` +
   residual, the mean relations, the output sign decoders and all gradients.  The same reference runs are stored as
   ``tests/golden/reference_models.pt`` (``tools/make_golden.py``) for machines without /root/reference.
2. ``tests/test_oracle_anchor.py``: the reference's known-answer test (tests/testGnnLightning.py:L214-216: MSE 6.33834 /
   RMSE 2.51761 / L1 2.31058 for the shipped checkpoint ``tests/test_models/epoch=48-val_MSE_loss=6.33834.ckpt`` on
   its 20 graphs), through fixtures extracted by ``tools/make_golden.py``; this one anchors the shim's GraphConv /
   HeteroConv / HeteroDictLinear semantics to numbers produced by the real torch_geometric.
3. Loss / metric literals and graph-template pins of the reference's tests (same file).
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence, Tuple

import torch
from torch import nn

EdgeType = Tuple[str, str, str]

MEAN_RELATIONS = ("gt", "gs", "center_bb")  # hgnn_k4.py:L107-120, hgnn_c2.py:L98-104


def edge_key(et: EdgeType) -> str:
    """PyG's ModuleDict key for an edge type: '<src___rel___dst>'."""
    return "<" + "___".join(et) + ">"


# ----------------------------------------------------------------------------------------
# group tables (yaml -> +-1 vectors)
# ----------------------------------------------------------------------------------------
def load_group_yaml(path: str) -> dict:
    import yaml
    with open(path, "r") as f:
        return yaml.safe_load(f)


def k4_tables(group: Optional[dict], with_feet: bool = True):
    """hgnn_k4.py:L37-94 / hgnn_k4_com.py:L37-82: order per local index = [e, gt, gs, gs*gt]."""
    one = torch.ones(3, dtype=torch.float64)

    def quad(key):
        if group is None:
            return torch.ones(12, dtype=torch.float64)
        r = group[key]  # IndexError / TypeError if the block is absent, like the reference
        gs = torch.tensor(r[0][:3], dtype=torch.float64)
        gt = torch.tensor(r[1][:3], dtype=torch.float64)
        return torch.cat((one, gt, gs, gs * gt))

    t = {"joint": quad("reflection_Q_js")}
    if with_feet:
        t["foot"] = quad("reflection_Q_fs")
    t["base_lin"] = quad("reflection_Q_bs_lin")
    t["base_ang"] = quad("reflection_Q_bs_ang")
    return t


def c2_tables(group: Optional[dict], with_feet: bool = True):
    """hgnn_c2.py:L42-83 / hgnn_c2_com.py: legs [e, e, gs, gs], bases [e, gs]."""
    one = torch.ones(3, dtype=torch.float64)

    def gs_of(key):
        if group is None:
            return one.clone()
        return torch.tensor(group[key][0][:3], dtype=torch.float64)

    js = gs_of("reflection_Q_js")
    t = {"joint": torch.cat((one, one, js, js))}
    if with_feet:
        fs = gs_of("reflection_Q_fs")
        t["foot"] = torch.cat((one, one, fs, fs))
    t["base_lin"] = torch.cat((one, gs_of("reflection_Q_bs_lin")))
    t["base_ang"] = torch.cat((one, gs_of("reflection_Q_bs_ang")))
    return t


# ----------------------------------------------------------------------------------------
# ReLU tap: lets a test record every pre-activation and force the sign decision of chosen elements
# ----------------------------------------------------------------------------------------
class ReluTap:
    """While installed (``with ReluTap() as tap``) every ReLU of the oracle goes through ``_relu``:
    ``record=True`` keeps the pre-activation tensors (in call order, with gradients retained);
    ``forced[i] = (flat_idx LongTensor, BoolTensor)`` overrides the (x > 0) decision of those elements of call i.
    Used by the parity tests to compare gradients under the ReLU sign pattern an fp32 implementation actually
    took where the fp64 pre-activation is numerically ambiguous (|x| below the implementation's forward error)."""

    def __init__(self, record: bool = False, forced=None):
        self.record = record
        self.forced = forced or {}
        self.pre: List[torch.Tensor] = []
        self.tags: List[tuple] = []      # ("enc", type) | ("conv", layer, type) | ("mlp", layer, type), one per recorded call
        self.ctx = None                  # tag of the next call made from inside an nn.Module (base_transform's ReLU)
        self.counter = 0

    def __enter__(self):
        global _TAP
        self._prev = _TAP
        _TAP = self
        return self

    def __exit__(self, *a):
        global _TAP
        _TAP = self._prev


_TAP: Optional[ReluTap] = None


def _relu(x: torch.Tensor, tag=None) -> torch.Tensor:
    tap = _TAP
    if tap is None:
        return torch.relu(x)
    i = tap.counter
    tap.counter += 1
    if tap.record:
        if x.requires_grad:
            x.retain_grad()
        tap.pre.append(x)
        tap.tags.append(tag if tag is not None else tap.ctx)
    mask = x > 0
    if i in tap.forced:
        idx, val = tap.forced[i]
        mask = mask.clone()
        mask.view(-1)[idx] = val
    return x * mask.to(x.dtype)


class _TapReLU(nn.Module):
    def forward(self, x):
        return _relu(x)


# ----------------------------------------------------------------------------------------
# building blocks
# ----------------------------------------------------------------------------------------
class _Lin(nn.Module):
    """weight [out, in] (+ bias [out]) with PyG's initialisation."""

    def __init__(self, in_f: int, out_f: int, bias: bool = True):
        super().__init__()
        self.weight = nn.Parameter(torch.empty(out_f, in_f))
        self.bias = nn.Parameter(torch.empty(out_f)) if bias else None
        self.reset_parameters()

    def reset_parameters(self):
        bound = 1.0 / math.sqrt(self.weight.shape[1]) if self.weight.shape[1] > 0 else 0.0
        with torch.no_grad():
            self.weight.uniform_(-bound, bound)
            if self.bias is not None:
                self.bias.uniform_(-bound, bound)

    def forward(self, x):
        y = x @ self.weight.t()
        return y if self.bias is None else y + self.bias


class _GraphConv(nn.Module):
    def __init__(self, h: int, mean: bool):
        super().__init__()
        self.mean = mean
        self.lin_rel = _Lin(h, h, bias=True)
        self.lin_root = _Lin(h, h, bias=False)

    def reset_parameters(self):
        self.lin_rel.reset_parameters()
        self.lin_root.reset_parameters()

    def forward(self, x_src, x_dst, edge_index):
        src, dst = edge_index[0], edge_index[1]
        agg = torch.zeros(x_dst.shape[0], x_src.shape[1], dtype=x_src.dtype, device=x_src.device)
        agg = agg.index_add(0, dst, x_src.index_select(0, src))
        if self.mean:
            deg = torch.bincount(dst, minlength=x_dst.shape[0]).clamp(min=1).to(agg.dtype)
            agg = agg / deg[:, None]
        return self.lin_rel(agg) + self.lin_root(x_dst)


class _HeteroConv(nn.Module):
    def __init__(self, h: int, edge_types: Sequence[EdgeType], mean_rels: Sequence[str]):
        super().__init__()
        self.edge_types = [tuple(e) for e in edge_types]
        self.convs = nn.ModuleDict({edge_key(e): _GraphConv(h, e[1] in mean_rels) for e in self.edge_types})

    def reset_parameters(self):
        for c in self.convs.values():
            c.reset_parameters()

    def forward(self, x_dict, edge_index_dict):
        out: Dict[str, torch.Tensor] = {}
        for et in self.edge_types:
            if et not in edge_index_dict:
                continue
            s, _, d = et
            o = self.convs[edge_key(et)](x_dict[s], x_dict[d], edge_index_dict[et])
            out[d] = o if d not in out else out[d] + o
        return out


class _Encoder(nn.Module):
    """HeteroDictLinear(-1, H, node_types): lazily sized per-type Linear."""

    def __init__(self, h: int, node_types: Sequence[str], in_dims: Optional[Dict[str, int]] = None):
        super().__init__()
        self.h = h
        self.node_types = list(node_types)
        self.lins = nn.ModuleDict()
        if in_dims is not None:
            self.materialize(in_dims)

    def materialize(self, in_dims: Dict[str, int]):
        for t in self.node_types:
            if t not in self.lins:
                self.lins[t] = _Lin(int(in_dims[t]), self.h, bias=True)

    def reset_parameters(self):
        for l in self.lins.values():
            l.reset_parameters()

    def forward(self, x_dict):
        if len(self.lins) < len(self.node_types):
            self.materialize({t: x_dict[t].shape[1] for t in self.node_types})
            self.to(next(iter(x_dict.values())).dtype)
        return {t: self.lins[t](x_dict[t]) for t in self.node_types if t in x_dict}


# ----------------------------------------------------------------------------------------
# the models
# ----------------------------------------------------------------------------------------
class _HGNNBase(nn.Module):
    """Common stack: encoder -> L x HeteroConv (+ MS post-processing) -> decoder."""

    morph_sym = False       # base_transform + residual (MS-HGNN) vs plain ReLU (MI-HGNN)
    decode_type = "foot"

    def __init__(self, hidden_channels, num_layers, data_metadata, out_channels, mean_rels=(),
                 in_dims=None):
        super().__init__()
        node_types, edge_types = data_metadata
        self.hidden_channels = hidden_channels
        self.num_layers = num_layers
        self.node_types = list(node_types)
        self.edge_types = [tuple(e) for e in edge_types]
        self.encoder = _Encoder(hidden_channels, node_types, in_dims)
        self.convs = nn.ModuleList(
            [_HeteroConv(hidden_channels, self.edge_types, mean_rels) for _ in range(num_layers)])
        if self.morph_sym:
            # one MLP shared by every layer: hgnn_k4.py:L133-137
            self.base_transform = nn.Sequential(
                _TorchLinear(hidden_channels, hidden_channels), _TapReLU(),
                _TorchLinear(hidden_channels, hidden_channels))
        self.decoder = _Lin(hidden_channels, out_channels, bias=True)

    def reset_parameters(self):
        self.encoder.reset_parameters()
        for c in self.convs:
            c.reset_parameters()
        if self.morph_sym:
            self.base_transform[0].reset_parameters()
            self.base_transform[2].reset_parameters()
        self.decoder.reset_parameters()

    # hooks -----------------------------------------------------------------------------
    def input_signs(self, x_dict):
        return x_dict

    def output_signs(self, out):
        return out

    def embed(self, x_dict, edge_index_dict):
        x_dict = self.input_signs(dict(x_dict))   # never mutate the caller's dict
        h = {k: _relu(v, ("enc", k)) for k, v in self.encoder(x_dict).items()}
        for l, conv in enumerate(self.convs):
            c = conv(h, edge_index_dict)
            if self.morph_sym:
                # hgnn_k4.py:L175-186: base -> shared MLP (no ReLU around it), others -> ReLU, then residual
                if _TAP is not None:
                    _TAP.ctx = ("mlp", l, "base")
                n = {k: (self.base_transform(v) if k == "base" else _relu(v, ("conv", l, k))) for k, v in c.items()}
                h = {k: (n[k] + h[k] if (k in h and h[k].shape == n[k].shape) else n[k]) for k in n}
            else:
                h = {k: _relu(v, ("conv", l, k)) for k, v in c.items()}   # hgnn.py:L60-62
        return h

    def forward(self, x_dict, edge_index_dict):
        h = self.embed(x_dict, edge_index_dict)
        return self.output_signs(self.decoder(h[self.decode_type]))


class _TorchLinear(nn.Linear):
    """nn.Linear (torch default init) - base_transform uses torch.nn.Linear, not PyG Linear."""


def _blockwise(x, n_nodes, signs_a, signs_b, T):
    """rows [B*n_nodes, 2*3*T] viewed [B, n, 2, 3, T]; first half * signs_a[3n+d], second * signs_b."""
    B = x.shape[0] // n_nodes
    v = x.reshape(B, n_nodes, 2, 3, T)
    s = torch.stack((signs_a.reshape(n_nodes, 3), signs_b.reshape(n_nodes, 3)), dim=1)  # [n, 2, 3]
    return (v * s.to(x.dtype).to(x.device)[None, :, :, :, None]).reshape(x.shape)


class GRF_HGNN(_HGNNBase):
    """MI-HGNN baseline, hgnn.py:L5-63."""

    def __init__(self, hidden_channels, num_layers, data_metadata, regression=True,
                 activation_fn=None, grf_dimension=1, in_dims=None):
        self.regression = regression
        self.grf_dimension = grf_dimension
        if regression and grf_dimension == 1:
            c = 1
        elif regression and grf_dimension == 3:
            c = 3
        else:
            c = 2
        self.out_channels_per_foot = c
        super().__init__(hidden_channels, num_layers, data_metadata, c, (), in_dims)


class COM_HGNN(_HGNNBase):
    """hgnn.py:L66-118."""
    decode_type = "base"

    def __init__(self, hidden_channels, num_layers, data_metadata, regression=True,
                 activation_fn=None, com_dimension=6, in_dims=None):
        self.regression = regression
        self.num_bases = 1
        self.num_dimensions_per_base = com_dimension
        super().__init__(hidden_channels, num_layers, data_metadata, com_dimension, (), in_dims)


class COM_HGNN_S4(COM_HGNN):
    """hgnn_s4_com.py:L6-71 (same arithmetic as COM_HGNN with 6 outputs)."""

    def __init__(self, hidden_channels, num_layers, data_metadata, regression=True,
                 activation_fn=None, symmetry_mode=None, group_operator_path=None, in_dims=None):
        super().__init__(hidden_channels, num_layers, data_metadata, regression, activation_fn, 6, in_dims)


class GRF_HGNN_K4(_HGNNBase):
    """hgnn_k4.py:L10-196."""
    morph_sym = True
    num_timesteps = 150
    num_legs = 4
    num_bases = 4
    num_joints = 12

    def __init__(self, hidden_channels, num_layers, data_metadata, regression=True, activation_fn=None,
                 symmetry_mode=None, group_operator_path=None, in_dims=None):
        self.regression = regression
        self.out_channels_per_foot = 1 if regression else 2
        group = load_group_yaml(group_operator_path) if (symmetry_mode and group_operator_path) else None
        t = k4_tables(group)
        super().__init__(hidden_channels, num_layers, data_metadata, self.out_channels_per_foot,
                         MEAN_RELATIONS, in_dims)
        self.joints_linear_weights = t["joint"]
        self.feet_linear_weights = t["foot"]
        self.base_coefficients_lin = t["base_lin"]
        self.base_coefficients_ang = t["base_ang"]

    def input_signs(self, x):
        # hgnn_k4.py:L198-237: whole joint row * sign[j]; foot/base rows blockwise
        j = x["joint"]
        B = j.shape[0] // self.num_joints
        x["joint"] = (j.reshape(B, self.num_joints, -1)
                      * self.joints_linear_weights.to(j.dtype).to(j.device)[None, :, None]).reshape(j.shape)
        x["foot"] = _blockwise(x["foot"], self.num_legs, self.feet_linear_weights,
                               self.feet_linear_weights, self.num_timesteps)
        x["base"] = _blockwise(x["base"], self.num_bases, self.base_coefficients_lin,
                               self.base_coefficients_ang, self.num_timesteps)
        return x


class GRF_HGNN_C2(_HGNNBase):
    """hgnn_c2.py:L10-189."""
    morph_sym = True
    num_timesteps = 150
    num_legs = 4
    num_bases = 2
    num_joints = 12

    def __init__(self, hidden_channels, num_layers, data_metadata, regression=True, activation_fn=None,
                 symmetry_mode=None, group_operator_path=None, grf_dimension=3, in_dims=None):
        self.regression = regression
        self.grf_dimension = grf_dimension
        if regression and grf_dimension == 1:
            c = 1
        elif regression and grf_dimension == 3:
            c = 3
        else:
            c = 2
        self.out_channels_per_foot = c
        group = load_group_yaml(group_operator_path) if (symmetry_mode and group_operator_path) else None
        t = c2_tables(group)
        super().__init__(hidden_channels, num_layers, data_metadata, c, MEAN_RELATIONS, in_dims)
        self.joints_linear_weights = t["joint"]
        self.feet_linear_weights = t["foot"]
        self.base_coefficients_lin = t["base_lin"]
        self.base_coefficients_ang = t["base_ang"]

    def input_signs(self, x):
        j = x["joint"]
        B = j.shape[0] // self.num_joints
        x["joint"] = (j.reshape(B, self.num_joints, -1)
                      * self.joints_linear_weights.to(j.dtype).to(j.device)[None, :, None]).reshape(j.shape)
        if not self.regression:   # hgnn_c2.py:L206: A1 foot features are a constant, untouched
            x["foot"] = _blockwise(x["foot"], self.num_legs, self.feet_linear_weights,
                                   self.feet_linear_weights, self.num_timesteps)
        x["base"] = _blockwise(x["base"], self.num_bases, self.base_coefficients_lin,
                               self.base_coefficients_ang, self.num_timesteps)
        return x

    def output_signs(self, out):
        if self.regression and self.grf_dimension == 3:    # ms_foot_decoder, hgnn_c2.py:L184-189
            out = out.reshape(-1, self.num_legs * 3)
            return out * self.feet_linear_weights.to(out.dtype).to(out.device)
        return out


class _COM_SYM(_HGNNBase):
    morph_sym = True
    decode_type = "base"
    num_joints = 12
    num_dimensions_per_base = 6

    def input_signs(self, x):
        # hgnn_k4_com.py:L167-177: joints only, [12, 2] view
        j = x["joint"]
        B = j.shape[0] // self.num_joints
        x["joint"] = (j.reshape(B, self.num_joints, -1)
                      * self.joints_linear_weights.to(j.dtype).to(j.device)[None, :, None]).reshape(j.shape)
        return x

    def output_signs(self, out):
        # morphological_symmetry_decoder, hgnn_k4_com.py:L157-165 -> [B, n_base, 6]
        B = out.shape[0] // self.num_bases
        o = out.reshape(B, self.num_bases, 6)
        s = torch.cat((self.base_coefficients_lin.reshape(self.num_bases, 3),
                       self.base_coefficients_ang.reshape(self.num_bases, 3)), dim=1)
        return o * s.to(out.dtype).to(out.device)[None]


class COM_HGNN_K4(_COM_SYM):
    """hgnn_k4_com.py:L10-177."""
    num_bases = 4

    def __init__(self, hidden_channels, num_layers, data_metadata, regression=True, activation_fn=None,
                 symmetry_mode=None, group_operator_path=None, in_dims=None):
        self.regression = regression
        group = load_group_yaml(group_operator_path) if (symmetry_mode and group_operator_path) else None
        t = k4_tables(group, with_feet=False)
        super().__init__(hidden_channels, num_layers, data_metadata, 6, MEAN_RELATIONS, in_dims)
        self.joints_linear_weights = t["joint"]
        self.base_coefficients_lin = t["base_lin"]
        self.base_coefficients_ang = t["base_ang"]


class COM_HGNN_C2(_COM_SYM):
    """hgnn_c2_com.py:L10-172."""
    num_bases = 2

    def __init__(self, hidden_channels, num_layers, data_metadata, regression=True, activation_fn=None,
                 symmetry_mode=None, group_operator_path=None, in_dims=None):
        self.regression = regression
        group = load_group_yaml(group_operator_path) if (symmetry_mode and group_operator_path) else None
        t = c2_tables(group, with_feet=False)
        super().__init__(hidden_channels, num_layers, data_metadata, 6, MEAN_RELATIONS, in_dims)
        self.joints_linear_weights = t["joint"]
        self.base_coefficients_lin = t["base_lin"]
        self.base_coefficients_ang = t["base_ang"]


# ----------------------------------------------------------------------------------------
# loss heads (L4 part of the hot path)
# ----------------------------------------------------------------------------------------
def contact_ce_loss(y_pred: torch.Tensor, y: torch.Tensor) -> torch.Tensor:
    """gnnLightning.py:L131-139 + customMetrics.py:L6-25.

    y_pred [B, 8] logits (no-contact, contact per foot), y [B, 4] in {0, 1}.
    CE(sum) over the 4B two-way rows, rounded through float32 (``.float()``), divided by 4B.
    """
    B = y_pred.shape[0]
    logits = y_pred.reshape(B * 4, 2)
    tgt = y.long().flatten()
    s = torch.nn.functional.cross_entropy(logits, tgt, reduction="sum")
    return s.float() / float(B * 4)


def contact_ce_loss_exact(y_pred: torch.Tensor, y: torch.Tensor) -> torch.Tensor:
    """Same value without the float32 rounding (for gradient checks)."""
    B = y_pred.shape[0]
    return torch.nn.functional.cross_entropy(y_pred.reshape(B * 4, 2), y.long().flatten(), reduction="mean")


def mse_loss(y_pred: torch.Tensor, y: torch.Tensor) -> torch.Tensor:
    """gnnLightning.py:L125-130, L633-639, gnnLightning_com.py:L96-121: mean over all elements."""
    d = y_pred.flatten() - y.flatten()
    return (d * d).mean()


def rmse_loss(y_pred, y):
    return torch.sqrt(mse_loss(y_pred, y))


def l1_loss(y_pred, y):
    return (y_pred.flatten() - y.flatten()).abs().mean()


def sixteen_class(y_pred_logits: torch.Tensor, y: torch.Tensor):
    """gnnLightning.py:L285-348: returns (pred class [B], true class [B])."""
    B = y_pred_logits.shape[0]
    p1 = torch.softmax(y_pred_logits.reshape(B * 4, 2), dim=1)[:, 1].reshape(B, 4)
    probs = torch.zeros(B, 16, dtype=p1.dtype)
    for j in range(16):
        f = []
        for k, div in enumerate((8, 4, 2, 1)):
            on = (j // div) % 2
            f.append(p1[:, k] if on else 1 - p1[:, k])
        probs[:, j] = (f[0] * f[1]) * (f[2] * f[3])
    y16 = (y[:, 0] * 8 + y[:, 1] * 4 + y[:, 2] * 2 + y[:, 3]).long()
    return probs.argmax(dim=1), y16


def binary_f1(pred: torch.Tensor, target: torch.Tensor) -> float:
    """customMetrics.py:L27-54."""
    pred = pred.long(); target = target.long()
    tp = int(((pred == 1) & (target == 1)).sum()); fp = int(((pred == 1) & (target == 0)).sum())
    fn = int(((pred == 0) & (target == 1)).sum())
    if tp + fp == 0 or tp + fn == 0:
        return 0.0
    p = tp / (tp + fp); r = tp / (tp + fn)
    return 0.0 if p + r == 0 else 2 * p * r / (p + r)


# ----------------------------------------------------------------------------------------
# batching (PyG Batch.from_data_list layout)
# ----------------------------------------------------------------------------------------
def tile_edge_index(template: torch.Tensor, n_src: int, n_dst: int, B: int) -> torch.Tensor:
    """[2, E] per-graph edges -> [2, E*B] batched: graph g's block offset by g*n_src / g*n_dst."""
    E = template.shape[1]
    g = torch.arange(B, dtype=torch.long).repeat_interleave(E)
    t = template.long().repeat(1, B)
    return torch.stack((t[0] + g * n_src, t[1] + g * n_dst))


# ----------------------------------------------------------------------------------------
# reading reference Lightning checkpoints without torch_geometric / lightning installed
# ----------------------------------------------------------------------------------------
class _StubMeta(type):
    def __getattr__(cls, name):
        raise AttributeError(name)


def _make_stub(mod, name):
    class Stub:
        _stub_of = f"{mod}.{name}"

        def __init__(self, *a, **k):
            self._args = a; self._kwargs = k; self._state = None

        def __setstate__(self, state):
            self._state = state

        def __call__(self, *a, **k):
            return _make_stub(mod, name + "()")(*a, **k)
    Stub.__name__ = name
    return Stub


class _PermissiveUnpickler:
    """pickle_module for torch.load: unknown classes become inert stubs."""
    import pickle as _p
    __name__ = "oracle_permissive_pickle"

    class Unpickler(_p.Unpickler):
        def find_class(self, mod, name):
            import importlib
            try:
                return getattr(importlib.import_module(mod), name)
            except Exception:
                return _make_stub(mod, name)

    @staticmethod
    def load(f, **kw):
        return _PermissiveUnpickler.Unpickler(f, **kw).load()


def load_reference_checkpoint(path: str):
    """Returns (state_dict without 'model.' prefix, hyper_parameters, batch dict or None)."""
    ck = torch.load(path, map_location="cpu", weights_only=False, pickle_module=_PermissiveUnpickler)
    sd = {k[len("model."):]: v for k, v in ck["state_dict"].items() if k.startswith("model.")}
    hp = ck.get("hyper_parameters", {})
    batch = None
    db = hp.get("dummy_batch", None)
    if db is not None and getattr(db, "_state", None) is not None:
        st = db._state
        x = {t: s._state["_mapping"]["x"] for t, s in st["_node_store_dict"].items()}
        ei = {tuple(e): s._state["_mapping"]["edge_index"] for e, s in st["_edge_store_dict"].items()}
        y = st["_global_store"]._state["_mapping"]["y"]
        batch = {"x": x, "edge_index": ei, "y": y}
    return sd, hp, batch
