"""Builds libgf2_b200.so (all CUDA kernels + the C ABI) for sm_100a, in-tree, with plain nvcc."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT = os.path.join(HERE, "libgf2_b200.so")
SOURCES = ["gf2_solver.cu", "gf2_tracker.cu", "gf2_lio.cu"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-Xcompiler", "-fPIC",
              "--expt-relaxed-constexpr", "-Xptxas", "-v", "-shared", "-lcudart", "-ldl"]


def needs_build():
    if not os.path.exists(OUT):
        return True
    t = os.path.getmtime(OUT)
    for root in (CSRC, os.path.join(HERE, "host"), os.path.join(HERE, "..", "include")):
        for f in os.listdir(root):
            if os.path.getmtime(os.path.join(root, f)) > t:
                return True
    return False


def build(force=False, verbose=False):
    if not force and not needs_build():
        return OUT
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    extra = os.environ.get("GF2_EXTRA_NVCC_FLAGS", "").split()  # debug builds only (e.g. -DGF2_PHASE_CLOCKS)
    cmd = [nvcc] + NVCC_FLAGS + extra + ["-o", OUT] + [os.path.join(CSRC, s) for s in SOURCES]
    res = subprocess.run(cmd, capture_output=True, text=True)
    log = res.stdout + res.stderr
    with open(os.path.join(HERE, "build.log"), "w") as f:
        f.write(" ".join(cmd) + "\n" + log)
    if verbose or res.returncode != 0:
        print(log)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed building libgf2_b200.so")
    # host-side C++ mirror of the reference classes (Estimator / FeatureManager / FeatureTracker glue) over the C ABI
    host_out = os.path.join(HERE, "libgf2_host.so")
    cmd = [os.environ.get("CXX", "g++"), "-O2", "-std=c++17", "-fPIC", "-shared", "-Wall", "-o", host_out, os.path.join(HERE, "host", "gf2_host.cpp"),
           "-L" + HERE, "-lgf2_b200", "-Wl,-rpath,$ORIGIN", "-Wl,--exclude-libs,ALL"]
    res = subprocess.run(cmd, capture_output=True, text=True)
    with open(os.path.join(HERE, "build.log"), "a") as f:
        f.write(" ".join(cmd) + "\n" + res.stdout + res.stderr)
    if verbose or res.returncode != 0:
        print(res.stdout + res.stderr)
    if res.returncode != 0:
        raise RuntimeError("g++ failed building libgf2_host.so")
    return OUT


if __name__ == "__main__":
    build(force="--force" in sys.argv, verbose=True)
