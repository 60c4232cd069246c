"""TEST INFRASTRUCTURE ONLY - a minimal stand-in for ``torch_geometric`` 2.5.0 (pinned by the reference's
pyproject.toml, absent from this image, not installable offline).

It exists for ONE purpose: to let ``tests/test_reference_pin.py`` and ``tools/make_golden.py`` import the reference's
UNMODIFIED model files (``/root/reference/src/ms_hgnn/lightning_py/hgnn*.py``; their only third-party imports are
``Linear, HeteroConv, HeteroDictLinear, GraphConv`` from ``torch_geometric.nn``) and run them, so that the oracle
(``oracle/mshgnn_oracle.py``) is pinned against the reference's own model code, gradients included.  Only the four
classes and only the call signatures the reference uses are provided; see ``nn/__init__.py``.
"""
__version__ = "2.5.0+shim"
