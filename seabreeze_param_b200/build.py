"""In-tree build of the CUDA C-ABI library (sm_100a only, no JIT cache, no CPU fallback)."""
import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
SRC = os.path.join(HERE, "csrc", "sbd_cuda.cu")
OUT_DIR = os.path.join(HERE, "_lib")
OUT = os.path.join(OUT_DIR, "libsbd_b200.so")
DEPS = [SRC, os.path.join(HERE, "csrc", "sbd_win8.cuh"), os.path.join(HERE, "csrc", "sbd_um.cuh"), os.path.join(HERE, "csrc", "sbd_f64.cuh"), os.path.join(ROOT, "include", "sbd.h"),
        os.path.join(ROOT, "include", "sbd_math.h")]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-fmad=false",            # reference arithmetic has separate mul/add; sbd_math.h relies on it
    "-Xcompiler", "-fPIC", "-shared",
    "-I" + os.path.join(ROOT, "include"),
]


def needs_build():
    if not os.path.exists(OUT):
        return True
    t = os.path.getmtime(OUT)
    return any(os.path.getmtime(d) > t for d in DEPS)


def build(force=False, verbose=False):
    """Compile csrc/sbd_cuda.cu -> _lib/libsbd_b200.so with nvcc (cross-compiles without a GPU)."""
    if not force and not needs_build():
        return OUT
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        raise RuntimeError("nvcc not found: cannot build the sm_100a sea-breeze kernels")
    os.makedirs(OUT_DIR, exist_ok=True)
    cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + [SRC, "-o", OUT]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + res.stdout + res.stderr)
    if verbose:
        print(res.stderr)
    return OUT


HARNESS_SRC = os.path.join(HERE, "csrc", "dummy_model_harness.cpp")
HARNESS = os.path.join(OUT_DIR, "dummy_model_harness")


def build_harness(force=False):
    """g++ build of the dummy_model call-sequence harness against the C-ABI library."""
    build(force=False)
    if not force and os.path.exists(HARNESS) and os.path.getmtime(HARNESS) >= max(
            os.path.getmtime(HARNESS_SRC), os.path.getmtime(OUT)):
        return HARNESS
    cmd = ["g++", "-O2", "-std=c++17", "-I" + os.path.join(ROOT, "include"), HARNESS_SRC, "-o", HARNESS,
           "-L" + OUT_DIR, "-lsbd_b200", "-Wl,-rpath,$ORIGIN"]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("harness build failed:\n" + res.stdout + res.stderr)
    return HARNESS


if __name__ == "__main__":
    import sys
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
    print(build_harness(force="--force" in sys.argv))
