"""Probe: pinned / pageable H2D bandwidth and zero-copy read bandwidth on this box."""
import time, torch, numpy as np
n = 256 * 1024 * 1024 // 4
d = torch.empty(n, dtype=torch.float32, device="cuda")
hp = torch.empty(n, dtype=torch.float32).pin_memory()
hg = torch.empty(n, dtype=torch.float32)
for name, h in (("pinned", hp), ("pageable", hg)):
    for _ in range(2):
        d.copy_(h, non_blocking=True); torch.cuda.synchronize()
    t = time.perf_counter()
    for _ in range(5):
        d.copy_(h, non_blocking=True)
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t) / 5
    print("H2D %s: %.1f GB/s" % (name, n * 4 / dt / 1e9))
t = time.perf_counter()
for _ in range(5):
    hp.copy_(d, non_blocking=True)
torch.cuda.synchronize()
print("D2H pinned: %.1f GB/s" % (n * 4 * 5 / (time.perf_counter() - t) / 1e9))
import subprocess
print(subprocess.run("nvidia-smi -q | grep -A6 'GPU Link Info' | head -12; nproc; free -g | head -2", shell=True, capture_output=True, text=True).stdout)
