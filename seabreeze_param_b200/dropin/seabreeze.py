"""Module named ``seabreeze``: put this directory on PYTHONPATH and the reference's
python_wrapper/seabreezediag/__init__.py (`from seabreeze import diag as dg, get_edges, get_dist`)
runs unchanged on the CUDA C-ABI."""
from seabreeze_param_b200.seabreeze import diag, get_edges, get_dist, sigmoid, get_threads  # noqa: F401
