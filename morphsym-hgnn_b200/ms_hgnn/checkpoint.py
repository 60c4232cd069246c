"""Reading the reference's Lightning checkpoints without lightning / torch_geometric installed.

``save_hyperparameters()`` in the reference pickles the whole dummy ``HeteroDataBatch``
(gnnLightning.py:L494), so its ``.ckpt`` files reference torch_geometric classes.  The loader maps
every un-importable class to an inert stub and digs the tensors out of the stub state.
"""
from __future__ import annotations

import importlib
import pickle

import torch


def _stub(mod: str, name: str):
    class Stub:
        _stub_of = f"{mod}.{name}"

        def __init__(self, *a, **k):
            self._args, self._kwargs, self._state = a, k, None

        def __setstate__(self, state):
            self._state = state

        def __call__(self, *a, **k):
            return _stub(mod, name + "()")(*a, **k)

    Stub.__name__ = name
    return Stub


class _Pickle:
    __name__ = "ms_hgnn_permissive_pickle"

    class Unpickler(pickle.Unpickler):
        def find_class(self, mod, name):
            try:
                return getattr(importlib.import_module(mod), name)
            except Exception:
                return _stub(mod, name)

    @staticmethod
    def load(f, **kw):
        return _Pickle.Unpickler(f, **kw).load()


def load_checkpoint(path: str, map_location="cpu") -> dict:
    """torch.load that tolerates missing third-party classes."""
    return torch.load(path, map_location=map_location, weights_only=False, pickle_module=_Pickle)


def model_state_dict(ckpt: dict, prefix: str = "model.") -> dict:
    return {k[len(prefix):]: v for k, v in ckpt["state_dict"].items() if k.startswith(prefix)}


def embedded_batch(ckpt: dict):
    """(x_dict, edge_index_dict, y) of the dummy batch a reference checkpoint carries, or None."""
    db = ckpt.get("hyper_parameters", {}).get("dummy_batch", None)
    st = getattr(db, "_state", None)
    if st is None:
        return None
    x = {t: s._state["_mapping"]["x"] for t, s in st["_node_store_dict"].items()}
    ei = {tuple(e): s._state["_mapping"]["edge_index"] for e, s in st["_edge_store_dict"].items()}
    y = st["_global_store"]._state["_mapping"]["y"]
    return x, ei, y
