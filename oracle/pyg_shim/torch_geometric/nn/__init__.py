"""The four ``torch_geometric.nn`` classes the reference's model files import, restated from the published
semantics of torch_geometric 2.5.0 (TEST INFRASTRUCTURE ONLY, see the package docstring).

* ``Linear(in, out, bias=True)`` (torch_geometric/nn/dense/linear.py): weight ``[out, in]``; ``in = -1`` defers the
  allocation to the first forward; ``reset_parameters`` = kaiming_uniform(a=sqrt(5)) on the weight (== U(+-1/sqrt(in)))
  and U(+-1/sqrt(in)) on the bias; ``forward = F.linear``.
* ``HeteroDictLinear(in, out, types)`` (same file): one independent ``Linear`` per node type in ``self.lins[type]``;
  ``forward(x_dict)`` maps every key of ``x_dict`` that has a ``Linear`` through it.
* ``GraphConv(in, out, aggr='add', bias=True)`` (torch_geometric/nn/conv/graph_conv.py):
  ``out_i = lin_rel(AGG_{j -> i} x_j) + lin_root(x_i)``; ``lin_rel`` carries the bias, ``lin_root`` has none; for a
  bipartite input ``(x_src, x_dst)`` the aggregate has ``x_dst.size(0)`` rows; ``mean`` divides the scatter-sum by
  ``clamp(in_degree, min=1)`` (torch_geometric.utils.scatter, reduce='mean').
* ``HeteroConv(convs, aggr='sum')`` (torch_geometric/nn/conv/hetero_conv.py): sub-modules live in a ModuleDict keyed
  ``'<src___rel___dst>'`` (torch_geometric/nn/module_dict.py); ``forward`` walks the edge types in insertion order,
  calls ``conv((x_src, x_dst), edge_index)`` (``conv(x, edge_index)`` when ``src == dst``), collects the outputs per
  destination type and reduces each list with ``torch.stack(xs, dim=0).sum(0)`` (``group``, aggr='sum').
"""
import math
from typing import Dict, Optional, Tuple, Union

import torch
import torch.nn.functional as F
from torch import Tensor, nn


class Linear(nn.Module):
    def __init__(self, in_channels: int, out_channels: int, bias: bool = True, weight_initializer: Optional[str] = None,
                 bias_initializer: Optional[str] = None):
        super().__init__()
        self.in_channels, self.out_channels = in_channels, out_channels
        self._has_bias = bias
        if in_channels > 0:
            self.weight = nn.Parameter(torch.empty(out_channels, in_channels))
        else:
            self.weight = nn.parameter.UninitializedParameter()  # lazy like PyG: materialised by the first forward
        self.bias = nn.Parameter(torch.empty(out_channels)) if bias else None
        self.reset_parameters()

    def reset_parameters(self):
        if isinstance(self.weight, nn.parameter.UninitializedParameter):
            return
        nn.init.kaiming_uniform_(self.weight, a=math.sqrt(5))
        if self.bias is not None:
            bound = 1.0 / math.sqrt(self.in_channels) if self.in_channels > 0 else 0.0
            nn.init.uniform_(self.bias, -bound, bound)

    def forward(self, x: Tensor) -> Tensor:
        if isinstance(self.weight, nn.parameter.UninitializedParameter):
            self.in_channels = x.size(-1)
            self.weight.materialize((self.out_channels, self.in_channels), device=x.device, dtype=x.dtype)
            self.reset_parameters()
        return F.linear(x, self.weight, self.bias)


class HeteroDictLinear(nn.Module):
    def __init__(self, in_channels: Union[int, Dict[str, int]], out_channels: int, types=None, **kwargs):
        super().__init__()
        if isinstance(in_channels, dict):
            self.types = list(in_channels.keys())
        else:
            self.types = list(types)
            in_channels = {t: in_channels for t in self.types}
        self.in_channels, self.out_channels = in_channels, out_channels
        self.lins = nn.ModuleDict({t: Linear(c, out_channels, **kwargs) for t, c in in_channels.items()})

    def reset_parameters(self):
        for lin in self.lins.values():
            lin.reset_parameters()

    def forward(self, x_dict: Dict[str, Tensor]) -> Dict[str, Tensor]:
        return {k: self.lins[k](x) for k, x in x_dict.items() if k in self.lins}


class GraphConv(nn.Module):
    def __init__(self, in_channels: Union[int, Tuple[int, int]], out_channels: int, aggr: str = "add", bias: bool = True, **kwargs):
        super().__init__()
        if aggr not in ("add", "sum", "mean"):
            raise NotImplementedError(f"shim GraphConv: aggr={aggr!r}")
        if isinstance(in_channels, int):
            in_channels = (in_channels, in_channels)
        self.aggr = aggr
        self.in_channels, self.out_channels = in_channels, out_channels
        self.lin_rel = Linear(in_channels[0], out_channels, bias=bias)
        self.lin_root = Linear(in_channels[1], out_channels, bias=False)

    def reset_parameters(self):
        self.lin_rel.reset_parameters()
        self.lin_root.reset_parameters()

    def forward(self, x, edge_index: Tensor, edge_weight=None, size=None) -> Tensor:
        if edge_weight is not None:
            raise NotImplementedError("shim GraphConv: edge_weight")
        if isinstance(x, Tensor):
            x = (x, x)
        x_src, x_dst = x
        n_dst = x_dst.size(0) if x_dst is not None else (size[1] if size is not None else x_src.size(0))
        msg = x_src.index_select(0, edge_index[0])
        out = torch.zeros(n_dst, x_src.size(-1), dtype=x_src.dtype, device=x_src.device).index_add_(0, edge_index[1], msg)
        if self.aggr == "mean":
            deg = torch.zeros(n_dst, dtype=x_src.dtype, device=x_src.device).index_add_(
                0, edge_index[1], torch.ones(edge_index.size(1), dtype=x_src.dtype, device=x_src.device))
            out = out / deg.clamp(min=1).unsqueeze(-1)
        out = self.lin_rel(out)
        if x_dst is not None:
            out = out + self.lin_root(x_dst)
        return out


def _key(edge_type) -> str:
    return "<" + "___".join(edge_type) + ">"


class HeteroConv(nn.Module):
    def __init__(self, convs: Dict[Tuple[str, str, str], nn.Module], aggr: Optional[str] = "sum"):
        super().__init__()
        if aggr != "sum":
            raise NotImplementedError(f"shim HeteroConv: aggr={aggr!r}")
        self._edge_types = [tuple(k) for k in convs.keys()]
        self.convs = nn.ModuleDict({_key(k): m for k, m in convs.items()})
        self.aggr = aggr

    def reset_parameters(self):
        for conv in self.convs.values():
            conv.reset_parameters()

    def forward(self, x_dict: Dict[str, Tensor], edge_index_dict) -> Dict[str, Tensor]:
        outs: Dict[str, list] = {}
        for et in self._edge_types:
            if et not in edge_index_dict:
                continue
            src, _, dst = et
            conv = self.convs[_key(et)]
            if src == dst:
                out = conv(x_dict[src], edge_index_dict[et])
            else:
                out = conv((x_dict[src], x_dict[dst]), edge_index_dict[et])
            outs.setdefault(dst, []).append(out)
        return {k: (v[0] if len(v) == 1 else torch.stack(v, dim=0).sum(0)) for k, v in outs.items()}
