"""Build libmocap_b200.so (sm_100a) in-tree with nvcc.  `python -m mocapv2_b200.build`.

The shared library is the product: a C-ABI (include/mocap_b200.h) over hand-written CUDA kernels.  It is built
in-tree (mocapv2_b200/libmocap_b200.so, git-ignored) so that it travels to the GPU box with the repo snapshot.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
SOURCES = ["api.cu", "detect_filter.cu", "detect_scan_tma.cu", "detect_cluster.cu", "detect_blobs.cu", "geometry.cu"]
LIB = os.path.join(HERE, "libmocap_b200.so")
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "--fmad=true", "-Xcompiler", "-fPIC,-O2", "-Xptxas", "-v", "-shared", "-cudart", "shared"]


def needs_build():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cu", ".cuh"))] + [os.path.join(HERE, "..", "include", "mocap_b200.h")]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    if not force and not needs_build():
        return LIB
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    cmd = [nvcc] + NVCC_FLAGS + ["-o", LIB] + [os.path.join(CSRC, s) for s in SOURCES]
    out = subprocess.run(cmd, capture_output=True, text=True)
    log = os.path.join(HERE, "csrc", "_ptxas.log")
    with open(log, "w") as f:
        f.write(out.stdout + out.stderr)
    if out.returncode != 0:
        sys.stderr.write(out.stdout + out.stderr)
        raise RuntimeError("nvcc failed building libmocap_b200.so")
    if verbose:
        print(out.stdout + out.stderr)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
