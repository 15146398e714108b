#!/usr/bin/env python
"""Offline "install" of the reference's pure-Python wrapper into the git-ignored baseline/_ref.

The reference's extension module cannot be built here (numpy.distutils is gone, there is no Fortran
compiler), so `pip install --target baseline/_ref /root/reference/python_wrapper` fails; what CAN be
installed is the wrapper package python_wrapper/seabreezediag/__init__.py, unmodified.  It is copied
(never committed: baseline/_ref is in .gitignore, but not in .gpurunignore, so it travels to the GPU
box) so that tests/test_gpu_r2.py::test_reference_python_wrapper_runs_on_the_dropin can run the
reference's own diag() time loop against the CUDA drop-in module named `seabreeze` on the box, where
/root/reference does not exist.  configdir.py (netCDF4 at import time) is not copied.
"""
import os
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = "/root/reference/python_wrapper/seabreezediag/__init__.py"
DST_DIR = os.path.join(ROOT, "baseline", "_ref", "seabreezediag")


def install():
    if not os.path.exists(SRC):
        return None
    os.makedirs(DST_DIR, exist_ok=True)
    dst = os.path.join(DST_DIR, "__init__.py")
    shutil.copyfile(SRC, dst)
    return dst


if __name__ == "__main__":
    r = install()
    print(r if r else "reference tree not present; nothing installed")
    sys.exit(0)
