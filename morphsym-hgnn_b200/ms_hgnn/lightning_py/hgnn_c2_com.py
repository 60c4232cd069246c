"""``COM_HGNN_C2`` on the B200-native engine (reference ``hgnn_c2_com.py:L10-172``): 2 bases + 12 joints -> ``[B, 2, 6]``."""
from torch import nn

from ..morphology import c2_sign_tables, load_group
from .hgnn_k4_com import _COM_SYM


class COM_HGNN_C2(_COM_SYM):
    fixed_nodes_per_graph = {"base": 2, "joint": 12}

    def __init__(self, hidden_channels: int, num_layers: int, data_metadata, regression: bool = True,
                 activation_fn=nn.ReLU(), symmetry_mode: str = None, group_operator_path: str = None, in_dims=None):
        group = load_group(group_operator_path) if (symmetry_mode and group_operator_path) else None
        self._setup(regression, c2_sign_tables(group, with_feet=False), 2)
        super().__init__(hidden_channels, num_layers, data_metadata, 6, activation_fn, in_dims)
