"""CPU-only checks of the drop-in boundary: the CUDA library builds for sm_100a, loads, exports
every symbol include/sbd.h declares, and refuses to compute without a GPU (no CPU fallback)."""
import ctypes
import os
import re
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    hdr = open(os.path.join(ROOT, "include", "sbd.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    return sorted(set(re.findall(r"\b(sbd_[a-z0-9_]+)\s*\(", hdr)))


def test_header_symbols_all_exported():
    from seabreeze_param_b200 import _capi
    lib = _capi.load()
    decl = _declared_symbols()
    assert len(decl) >= 20
    for s in decl:
        assert hasattr(lib, s), "include/sbd.h declares %s but the library does not export it" % s
    assert set(decl) == set(_capi.SYMBOLS)


def test_library_is_sm100a_and_in_tree():
    from seabreeze_param_b200 import _capi
    path = _capi.lib_path()
    assert path.startswith(ROOT)
    out = subprocess.run(["cuobjdump", "-lelf", path], capture_output=True, text=True).stdout
    assert "sm_100a" in out


def test_no_cpu_fallback():
    """Without a CUDA device every compute entry point fails with SBD_ECUDA (3)."""
    from seabreeze_param_b200 import _capi, seabreeze
    lib = _capi.load()
    if lib.sbd_device_count() > 0:
        pytest.skip("GPU present")
    a = np.zeros((8, 8), np.float32, order="F")
    with pytest.raises(_capi.SbdError):
        seabreeze.get_edges(a, a)
    assert seabreeze.get_threads() == 0


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "seabreeze_param_b200")
    for dp, _, fs in os.walk(pkg):
        for f in fs:
            if f.endswith((".py", ".cu", ".h", ".cuh", ".cpp")):
                txt = open(os.path.join(dp, f)).read()
                for pat in (r"^\s*(from|import)\s+oracle", r"libsbd_oracle", r"orc_[a-z_]+\s*\(", r"oracle[/.]"):
                    assert not re.search(pat, txt, flags=re.M), "%s uses the oracle (%s)" % (f, pat)


def test_reference_wrapper_imports_unchanged():
    """The reference's own python_wrapper/seabreezediag/__init__.py must import against our
    drop-in module named `seabreeze` (import only; compute needs the GPU)."""
    ref = "/root/reference/python_wrapper"
    if not os.path.isdir(ref):
        pytest.skip("reference tree not present on this machine")
    code = ("import sys; sys.path[:0]=[%r,%r,%r]; import seabreezediag as s; "
            "import seabreeze; assert s.dg is seabreeze.diag; print('ok')"
            % (os.path.join(ROOT, "seabreeze_param_b200", "dropin"), ROOT, ref))
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True)
    assert r.returncode == 0 and "ok" in r.stdout, r.stderr


def test_f2py_mirror_argument_checks():
    from seabreeze_param_b200 import seabreeze
    a = np.zeros((8, 6), np.float32)
    with pytest.raises(ValueError):
        seabreeze.get_edges(a, np.zeros((6, 8), np.float32))
    with pytest.raises(ValueError):
        seabreeze.get_dist(a, a, np.zeros(7), np.zeros(6))


def test_argument_validation_status_codes():
    """Status codes of include/sbd.h that do not need a device: 1 = bad dimensions (the UM hook's
    `error = 1`, UM/vn10.7/sea_breeze_diag.F90:198-201), 2 = invalid argument."""
    import ctypes as C
    from seabreeze_param_b200 import _capi
    lib = _capi.load()
    h = C.c_void_p()
    assert lib.sbd_plan_create(C.byref(h), 0, 0, 10, 5, 2, None) == 1        # nx < 1
    assert lib.sbd_plan_create(C.byref(h), 0, 10, 10, 0, 2, None) == 1       # nz < 1
    assert lib.sbd_plan_create(C.byref(h), 7, 10, 10, 5, 2, None) == 2       # unknown variant
    assert lib.sbd_plan_create(C.byref(h), 0, 10, 10, 5, 0, None) == 2       # generic needs a halo
    assert lib.sbd_plan_create(None, 0, 10, 10, 5, 2, None) == 2
    a = np.zeros((8, 8), np.float32)
    assert lib.sbd_get_edges_f32(a.ctypes.data, a.ctypes.data, a.ctypes.data, 4, 4, 0, 1, None) == 2   # halo < 1
    assert lib.sbd_get_edges_f32(None, a.ctypes.data, a.ctypes.data, 4, 4, 1, 1, None) == 2
    assert lib.sbd_py_get_dist_f32(a.ctypes.data, a.ctypes.data, a.ctypes.data, a.ctypes.data, 1, 8, 180.0,
                                   a.ctypes.data, None) == 1                                           # nlons < 2
    assert lib.sbd_plan_step(None, 1.0, 1, *([a.ctypes.data] * 8), None) == 2
    assert b"" != lib.sbd_last_error()
    p = _capi.default_params(0)
    assert abs(p.target_plev - 70000.0) < 1e-3 and p.missing == 0.0 and abs(p.gmma + 0.0060956) < 1e-9
    q = _capi.default_params(1)
    assert q.missing == np.float32(2.0e20)
