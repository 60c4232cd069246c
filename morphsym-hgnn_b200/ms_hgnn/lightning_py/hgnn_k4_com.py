"""``COM_HGNN_K4`` on the B200-native engine (reference ``hgnn_k4_com.py:L10-177``).

Nodes: 4 bases + 12 joints (no feet); decodes the base nodes to ``[B, 4, 6]`` with the lin/ang
reflections of ``morphological_symmetry_decoder`` (L157-165) applied in the decoder kernel.
"""
from torch import nn

from ..modules import NativeHGNN
from ..morphology import c2_sign_tables, k4_sign_tables, load_group, rowwise_signs
from .hgnn_k4 import MEAN_RELATIONS, _nontrivial


class _COM_SYM(NativeHGNN):
    morph_sym = True
    decode_node = "base"
    mean_relations = MEAN_RELATIONS

    def _setup(self, regression, tables, num_bases):
        self.regression = regression
        self.num_timesteps = 1
        self.num_legs = 4
        self.num_bases = num_bases
        self.num_joints = 12
        self.num_dimensions_per_base = 6
        self.joints_linear_weights = tables["joint"]
        self.base_coefficients_lin = tables["base_lin"]
        self.base_coefficients_ang = tables["base_ang"]

    def _in_sign(self, in_dims):
        return {"joint": _nontrivial(rowwise_signs(self.joints_linear_weights.tolist(), in_dims["joint"])), "base": None}

    def _out_sign(self):
        s = []
        for b in range(self.num_bases):
            s += [float(v) for v in self.base_coefficients_lin[3 * b:3 * b + 3].tolist()]
            s += [float(v) for v in self.base_coefficients_ang[3 * b:3 * b + 3].tolist()]
        return _nontrivial(s)

    def _finish(self, out, B):
        return out.view(B, self.num_bases, 6)


class COM_HGNN_K4(_COM_SYM):
    fixed_nodes_per_graph = {"base": 4, "joint": 12}

    def __init__(self, hidden_channels: int, num_layers: int, data_metadata, regression: bool = True,
                 activation_fn=nn.ReLU(), symmetry_mode: str = None, group_operator_path: str = None, in_dims=None):
        group = load_group(group_operator_path) if (symmetry_mode and group_operator_path) else None
        self._setup(regression, k4_sign_tables(group, with_feet=False), 4)
        super().__init__(hidden_channels, num_layers, data_metadata, 6, activation_fn, in_dims)
