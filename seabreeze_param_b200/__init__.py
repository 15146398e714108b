"""seabreeze_param_b200 -- B200-native (sm_100a CUDA) implementation of the grid-point
land-sea-breeze trigger of antarcticrainforest/seabreeze_param (Bergemann et al. 2017).

Only the hot path is here: get_edges, get_dist, sigmoid, seabreeze_diag/diag and
swap_bounds, behind the reference's own interfaces:

  seabreeze_param_b200.seabreeze      mirror of the f2py extension module `seabreeze`
  seabreeze_param_b200.seabreezediag  mirror of python_wrapper/seabreezediag/__init__.py (diag, c2f)
  seabreeze_param_b200.generic        mirror of the Fortran module sea_breeze_diag_mod + Plan API
  seabreeze_param_b200.halo           swap_bounds across GPUs (NCCL send/recv of row-block halos)
  seabreeze_param_b200.synth          synthetic fields of BASELINE.json's configs

All compute runs in csrc/sbd_cuda.cu through the C-ABI of include/sbd.h; there is no CPU fallback.
"""
from ._capi import SbdError, load, lib_path  # noqa: F401

__all__ = ["SbdError", "load", "lib_path"]
