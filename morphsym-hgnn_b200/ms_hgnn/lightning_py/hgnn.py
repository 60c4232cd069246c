"""MI-HGNN baselines on the B200-native engine.

Drop-in for the reference's ``src/ms_hgnn/lightning_py/hgnn.py`` (``GRF_HGNN`` L5-63, ``COM_HGNN``
L66-118): same constructor signatures, ``forward(x_dict, edge_index_dict)``, attributes and
state-dict keys.  One base node per graph, no sign tables, ``h <- relu(conv(h))`` per layer.
"""
from torch import nn

from ..modules import NativeHGNN


def _grf_channels(regression: bool, grf_dimension: int) -> int:
    if regression and grf_dimension == 1:
        return 1
    if regression and grf_dimension == 3:
        return 3
    return 2          # contact logits (no contact, contact)


class GRF_HGNN(NativeHGNN):
    morph_sym = False
    decode_node = "foot"

    def __init__(self, hidden_channels: int, num_layers: int, data_metadata, regression: bool = True,
                 activation_fn=nn.ReLU(), grf_dimension: int = 1, in_dims=None, nodes_per_graph=None):
        self.regression = regression
        self.grf_dimension = grf_dimension
        self.out_channels_per_foot = _grf_channels(regression, grf_dimension)
        super().__init__(hidden_channels, num_layers, data_metadata, self.out_channels_per_foot, activation_fn,
                         in_dims, nodes_per_graph)


class COM_HGNN(NativeHGNN):
    morph_sym = False
    decode_node = "base"

    def __init__(self, hidden_channels: int, num_layers: int, data_metadata, regression: bool = True,
                 activation_fn=nn.ReLU(), com_dimension: int = 6, in_dims=None, nodes_per_graph=None):
        self.regression = regression
        self.num_bases = 1
        self.num_dimensions_per_base = com_dimension
        super().__init__(hidden_channels, num_layers, data_metadata, com_dimension, activation_fn, in_dims, nodes_per_graph)
