"""``COM_HGNN_S4`` (reference ``hgnn_s4_com.py:L6-71``): MI-HGNN stack decoding the single base node to 6 values."""
from torch import nn

from ..modules import NativeHGNN


class COM_HGNN_S4(NativeHGNN):
    morph_sym = False
    decode_node = "base"

    def __init__(self, hidden_channels: int, num_layers: int, data_metadata, regression: bool = True,
                 activation_fn=nn.ReLU(), symmetry_mode: str = None, group_operator_path: str = None,
                 in_dims=None, nodes_per_graph=None):
        self.regression = regression
        self.num_bases = 1
        self.num_dimensions_per_base = 6
        super().__init__(hidden_channels, num_layers, data_metadata, 6, activation_fn, in_dims, nodes_per_graph)
