"""The oracle against golden vectors produced by the REFERENCE'S OWN Fortran source.

tests/golden/ref_*.npz hold inputs and outputs of python_wrapper/seabreezediag/{sobel,seabreeze_diag_python}.f90
and of generic/sea_breeze_diag.f90's sigmoid, executed unmodified by tests/golden/f90interp.py (an interpreter for
the Fortran subset they use: binary32, one IEEE operation per Fortran operation, platform libm) -- see
tests/golden/make_reference_golden.py.  This is the pin SURVEY.md section 8c says the reference itself does not
provide: the oracle (oracle/sbd_oracle.cpp, a hand restatement) must reproduce every value BIT FOR BIT, in its
libm mode (the same libm calls the interpreter made) and in its pinned-math mode (what the GPU runs).
Cells the reference never writes (last latitude row, seabreeze_diag_python.f90:165) are NaN in the fixtures and
are not compared.  CPU only."""
import os

import numpy as np
import pytest

from util import assert_bits_equal

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _load(name):
    return np.load(os.path.join(G, name))


def _cmp_written(ref, got, what):
    w = ~np.isnan(ref)
    assert w.any(), what
    assert_bits_equal(np.where(w, ref, 0).astype(np.float32), np.where(w, got, 0).astype(np.float32), what)


@pytest.mark.parametrize("mode", [0, 1], ids=["libm", "pinned"])
@pytest.mark.parametrize("name", ["ref_py_a.npz", "ref_py_b.npz"])
def test_oracle_reproduces_reference_time_loop(oracle, name, mode):
    d = _load(name)
    oracle.set_math(mode)
    try:
        land, z, sig, lon, lat, p = (d[k] for k in ("land", "z", "sig", "lon", "lat", "p"))
        kw = {k[3:]: float(d[k]) for k in d.files if k.startswith("kw_")}
        ws = np.zeros_like(z); wd = np.zeros_like(z); thc = np.zeros_like(z)
        nrec = d["theta"].shape[0]
        for r in range(nrec):
            coast = oracle.py_get_edges(land, d["ci"][r])
            assert_bits_equal(d["coast"][r], coast, "%s get_edges record %d" % (name, r))
            for literal in (True, False):
                cd = oracle.py_get_dist(coast, land, lon, lat, maxdist=float(d["dist_maxdist"]), literal=literal)
                assert_bits_equal(d["cdist"][r], cd, "%s get_dist record %d literal=%s" % (name, r, literal))
            for literal in (True, False):
                o, f = oracle.py_diag(1 + r, p, z, sig, d["theta"][r], d["v"][r], d["u"][r], cd, ws, wd, thc,
                                      literal=literal, **kw)
                for k, nm in enumerate(("sb_con", "t0", "windspeed", "winddir")):
                    _cmp_written(d["out"][r][k], o[k], "%s %s record %d literal=%s" % (name, nm, r, literal))
            ref = d["out"][r]
            thc, ws, wd = [np.where(np.isnan(ref[k]), prev, ref[k]).astype(np.float32)
                           for k, prev in ((1, thc), (2, ws), (3, wd))]
        # the trigger fired somewhere and the missing value is where the coast is far
        sb = d["out"][:, 0]
        assert ((sb != 0) & (sb < 1e19)).any() and (sb == np.float32(2.0e20)).any()
    finally:
        oracle.set_math(1)


def test_oracle_reproduces_reference_edges_on_ragged_grids(oracle):
    d = _load("ref_py_c.npz")
    for i in range(3):
        assert_bits_equal(d["coast%d" % i], oracle.py_get_edges(d["lsm%d" % i], d["ci%d" % i]), "ragged grid %d" % i)


@pytest.mark.parametrize("mode", [0, 1], ids=["libm", "pinned"])
def test_oracle_reproduces_reference_sigmoid(oracle, mode):
    d = _load("ref_sigmoid.npz")
    oracle.set_math(mode)
    try:
        sm, st = oracle.sigmoid(d["ary"])
        assert_bits_equal(d["sm"], sm, "sigmoid (python and generic source give the same bits)")
    finally:
        oracle.set_math(1)


def test_fixture_provenance():
    """Each fixture records the sha256 of the reference files that produced it; when the reference tree is present
    the files must still be those."""
    import hashlib
    ref = "/root/reference"
    if not os.path.isdir(ref):
        pytest.skip("reference tree not present on this machine")
    paths = {"sobel": "python_wrapper/seabreezediag/sobel.f90", "diag": "python_wrapper/seabreezediag/seabreeze_diag_python.f90",
             "generic": "generic/sea_breeze_diag.f90"}
    d = _load("ref_py_a.npz")
    for k, rel in paths.items():
        assert str(d["sha256_" + k]) == hashlib.sha256(open(os.path.join(ref, rel), "rb").read()).hexdigest(), k


def test_interpreter_flags_the_generic_variants_undefined_behaviour():
    """generic/sea_breeze_diag.f90 cannot be pinned the same way: executed literally, its seabreeze_diag reads the
    logical `found` before any assignment (G:139, G:191) -- SURVEY.md A.6.  The interpreter reports exactly that."""
    ref = "/root/reference/generic/sea_breeze_diag.f90"
    if not os.path.exists(ref):
        pytest.skip("reference tree not present on this machine")
    import sys
    sys.path.insert(0, G)
    import f90interp as F
    M = F.Machine(open(ref).read())
    nx, ny, nz = 6, 5, 3
    f = lambda *s: np.asfortranarray(np.full(s, 280.0, np.float32))
    args = [np.float32(1440.0), 1, f(nx, ny, nz), f(nx, ny, nz), f(nx, ny, nz), f(nx, ny), f(nx, ny) * 0, f(nx, ny),
            np.asfortranarray(np.arange(nx * ny, dtype=np.float32).reshape(nx, ny)), f(nx, ny), f(nx, ny), f(nx, ny), f(nx, ny)]
    with pytest.raises(F.FortranError, match="found"):
        M.call("seabreeze_diag", *args)
