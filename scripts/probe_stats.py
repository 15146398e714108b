"""Probe: duration of the sigmoid statistics (exact parallel vs sequential) for the config sizes."""
import sys, os, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from seabreeze_param_b200 import generic
rng = np.random.default_rng(0)
for ny, nx in ((769, 1024), (1800, 3600), (3000, 4000)):
    a = (rng.random((ny, nx)) * 300 * (rng.random((ny, nx)) > 0.7)).astype(np.float32)
    d = torch.from_numpy(a).cuda(); sm = torch.empty_like(d); st = np.zeros(4, np.float32)
    generic.sigmoid(d, sm, st); torch.cuda.synchronize()
    t0 = time.perf_counter(); generic.sigmoid(d, sm, st); torch.cuda.synchronize()
    print("%dx%d: sigmoid (2 stat passes + apply) %.3f ms  stats=%s" % (nx, ny, (time.perf_counter() - t0) * 1e3, st))
