"""Residual rate of the pinned transcendental functions (include/sbd_math.h) against the platform libm -- the
functions gfortran lowers the reference's exp / atan2 / sin / cos to (G:167, G:227, S:129-132, S:175-186).

Pinned and libm results are compared bit for bit on 4 million inputs per function drawn over the ranges the path
uses (exp: the logistic argument -std*(sigma-r); sin/cos: latitudes / longitude differences in radians and well
beyond; atan2: wind components).  The exhaustive census of round 1 found 2 of 2.24e9 (expf), 12 and 22 of 1.12e9
(sinf, cosf), 0 (atanf, atan2f) differing inputs on glibc 2.39; here: at most 4 per million, never more than 1 ulp.
tests/test_libm_census.py measures the same thing where it matters: on the outputs of the whole path."""
import ctypes as C
import os
import subprocess
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _build():
    src = ('#include <math.h>\n#include "sbd_math.h"\n'
           'extern "C" void both(const float* x, const float* y, int n, float* pe, float* ps, float* pc, float* pa,'
           ' float* le, float* ls, float* lc, float* la) { for (int i = 0; i < n; ++i) {'
           ' pe[i] = sbd_expf(x[i]); ps[i] = sbd_sinf(x[i]); pc[i] = sbd_cosf(x[i]); pa[i] = sbd_atan2f(y[i], x[i]);'
           ' le[i] = expf(x[i]); ls[i] = sinf(x[i]); lc[i] = cosf(x[i]); la[i] = atan2f(y[i], x[i]); } }')
    d = tempfile.mkdtemp()
    open(os.path.join(d, "t.cpp"), "w").write(src)
    subprocess.check_call(["g++", "-O2", "-ffp-contract=off", "-fno-builtin", "-shared", "-fPIC", "-I" + os.path.join(ROOT, "include"),
                           os.path.join(d, "t.cpp"), "-o", os.path.join(d, "t.so"), "-lm"])
    return C.CDLL(os.path.join(d, "t.so"))


def test_pinned_functions_track_platform_libm():
    L = _build()
    rng = np.random.default_rng(7)
    n = 4_000_000
    x = np.concatenate([rng.uniform(-90.0, 90.0, n // 2), rng.normal(0.0, 8.0, n // 4),
                        rng.uniform(-3.2, 3.2, n // 4)]).astype(np.float32)
    y = rng.normal(0.0, 12.0, n).astype(np.float32)
    outs = [np.empty(n, np.float32) for _ in range(8)]
    q = lambda a: a.ctypes.data_as(C.c_void_p)
    L.both(q(x), q(y), C.c_int(n), *[q(a) for a in outs])
    for k, nm in enumerate(("expf", "sinf", "cosf", "atan2f")):
        p, l = outs[k], outs[k + 4]
        fin = np.isfinite(l)
        assert np.array_equal(np.isfinite(p), fin), nm
        pb, lb = p[fin].view(np.int32).astype(np.int64), l[fin].view(np.int32).astype(np.int64)
        diff = np.abs(pb - lb)
        n_diff = int((diff != 0).sum())
        assert n_diff <= 4 * (n // 1_000_000), "%s: %d of %d inputs differ from libm" % (nm, n_diff, n)
        assert diff.max() <= 1, "%s: %d ulp" % (nm, diff.max())
