"""``GRF_HGNN_C2`` on the B200-native engine (reference ``hgnn_c2.py:L10-189``).

C2 morphology: 2 base nodes, 12 joints, 4 feet.  Regression with ``grf_dimension == 3`` returns
``[B, 12]`` with the foot reflection applied (``ms_foot_decoder``, L184-189); the sign flip of
the foot inputs is skipped for regression (L206).
"""
from torch import nn

from ..modules import NativeHGNN
from ..morphology import blockwise_signs, c2_sign_tables, load_group, rowwise_signs
from .hgnn_k4 import MEAN_RELATIONS, _nontrivial


class GRF_HGNN_C2(NativeHGNN):
    morph_sym = True
    decode_node = "foot"
    mean_relations = MEAN_RELATIONS
    fixed_nodes_per_graph = {"base": 2, "joint": 12, "foot": 4}

    def __init__(self, hidden_channels: int, num_layers: int, data_metadata, regression: bool = True,
                 activation_fn=nn.ReLU(), symmetry_mode: str = None, group_operator_path: str = None,
                 grf_dimension: int = 3, in_dims=None):
        self.regression = regression
        self.num_timesteps = 150
        self.num_legs = 4
        self.num_bases = 2
        self.num_joints = 12
        self.num_dimensions_per_foot = 3
        self.num_dimensions_per_base = 3
        self.num_variables_per_joint = 3 if regression else 2
        self.grf_dimension = grf_dimension
        group = load_group(group_operator_path) if (symmetry_mode and group_operator_path) else None
        t = c2_sign_tables(group)
        self.joints_linear_weights = t["joint"]
        self.feet_linear_weights = t["foot"]
        self.base_coefficients_lin = t["base_lin"]
        self.base_coefficients_ang = t["base_ang"]
        if regression and grf_dimension == 1:
            c = 1
        elif regression and grf_dimension == 3:
            c = 3
        else:
            c = 2
        self.out_channels_per_foot = c
        super().__init__(hidden_channels, num_layers, data_metadata, c, activation_fn, in_dims)

    def _in_sign(self, in_dims):
        T = self.num_timesteps
        if in_dims["joint"] != self.num_variables_per_joint * T:
            raise ValueError(f"x_dict['joint'] must be {self.num_variables_per_joint * T} wide, got {in_dims['joint']}")
        if in_dims["base"] != 6 * T:
            raise ValueError(f"x_dict['base'] must be {6 * T} wide, got {in_dims['base']}")
        out = {"joint": _nontrivial(rowwise_signs(self.joints_linear_weights.tolist(), in_dims["joint"])),
               "base": _nontrivial(blockwise_signs(2, self.base_coefficients_lin, self.base_coefficients_ang, T)),
               "foot": None}
        if not self.regression:
            if in_dims["foot"] != 6 * T:
                raise ValueError(f"x_dict['foot'] must be {6 * T} wide, got {in_dims['foot']}")
            out["foot"] = _nontrivial(blockwise_signs(4, self.feet_linear_weights, self.feet_linear_weights, T))
        return out

    def _out_sign(self):
        if self.regression and self.grf_dimension == 3:
            return _nontrivial([float(v) for v in self.feet_linear_weights.tolist()])
        return None

    def _finish(self, out, B):
        if self.regression and self.grf_dimension == 3:
            return out.view(B, self.num_legs * 3)      # ms_foot_decoder returns [B, 12]
        return out
