"""``GRF_HGNN_K4`` on the B200-native engine (reference ``hgnn_k4.py:L10-196``).

K4 morphology: 4 base nodes (one per group element e, gt, gs, gr), 12 joints, 4 feet.  The +-1
input tables of ``apply_symmetry`` (L198-237) are folded into the encoder's feature load, the
fixed edge template into constant gather tables.
"""
import torch
from torch import nn

from ..modules import NativeHGNN
from ..morphology import blockwise_signs, k4_sign_tables, load_group, rowwise_signs

MEAN_RELATIONS = ("gt", "gs", "center_bb")


def _nontrivial(v):
    return v if any(s != 1.0 for s in v) else None


class GRF_HGNN_K4(NativeHGNN):
    morph_sym = True
    decode_node = "foot"
    mean_relations = MEAN_RELATIONS
    fixed_nodes_per_graph = {"base": 4, "joint": 12, "foot": 4}

    def __init__(self, hidden_channels: int, num_layers: int, data_metadata, regression: bool = True,
                 activation_fn=nn.ReLU(), symmetry_mode: str = None, group_operator_path: str = None, in_dims=None):
        self.regression = regression
        self.num_timesteps = 150            # hard-coded in the reference (L28-35)
        self.num_legs = 4
        self.num_bases = 4
        self.num_joints = 12
        self.num_dimensions_per_foot = 3
        self.num_dimensions_per_base = 3
        group = load_group(group_operator_path) if (symmetry_mode and group_operator_path) else None
        t = k4_sign_tables(group)           # raises like the reference when a yaml block is missing
        self.joints_linear_weights = t["joint"]
        self.feet_linear_weights = t["foot"]
        self.base_coefficients_lin = t["base_lin"]
        self.base_coefficients_ang = t["base_ang"]
        self.out_channels_per_foot = 1 if regression else 2
        super().__init__(hidden_channels, num_layers, data_metadata, self.out_channels_per_foot, activation_fn, in_dims)

    def _in_sign(self, in_dims):
        T = self.num_timesteps
        for k in ("foot", "base"):
            if in_dims[k] != 6 * T:
                raise ValueError(f"x_dict['{k}'] must be {6 * T} wide (2 variables x 3 axes x {T} steps), got {in_dims[k]}")
        if in_dims["joint"] != 2 * T:
            raise ValueError(f"x_dict['joint'] must be {2 * T} wide, got {in_dims['joint']}")
        return {
            "joint": _nontrivial(rowwise_signs(self.joints_linear_weights.tolist(), in_dims["joint"])),
            "foot": _nontrivial(blockwise_signs(4, self.feet_linear_weights, self.feet_linear_weights, T)),
            "base": _nontrivial(blockwise_signs(4, self.base_coefficients_lin, self.base_coefficients_ang, T)),
        }
