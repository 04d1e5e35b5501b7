"""Build tests/emu/_build/libmocap_emu.so: the product's kernel sources compiled for the CPU (TEST INFRASTRUCTURE ONLY).

See cuda_emu.h.  Used by the GPU-less part of the test-suite to run the kernel logic of mocapv2_b200/csrc/*.cu
against the oracle; the mocapv2_b200 package never loads it.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
CSRC = os.path.join(REPO, "mocapv2_b200", "csrc")
OUT = os.path.join(HERE, "_build")
LIB = os.path.join(OUT, "libmocap_emu.so")
SOURCES = ["api.cu", "detect_filter.cu", "detect_scan_tma.cu", "detect_cluster.cu", "detect_blobs.cu", "geometry.cu"]


def build(force=False):
    extra = os.environ.get("MOCAP_EMU_FLAGS", "").split()            # e.g. -DBAYER_NW_MAX=4 (variant checks; forces a rebuild)
    force = force or bool(extra)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cu", ".cuh"))] + \
           [os.path.join(HERE, "cuda_emu.h"), os.path.join(HERE, "cuda_emu.cpp"),
            os.path.join(REPO, "include", "mocap_b200.h")]
    if not force and os.path.exists(LIB) and all(os.path.getmtime(d) <= os.path.getmtime(LIB) for d in deps):
        return LIB
    os.makedirs(OUT, exist_ok=True)
    objs = []
    for s in SOURCES:
        o = os.path.join(OUT, s + ".o")
        subprocess.check_call(["g++", "-x", "c++", "-std=c++20", "-O1", "-g", "-ffp-contract=off", "-fPIC", "-DMOCAP_EMU",
                               "-Wno-attributes", "-I", HERE] + extra + ["-c", os.path.join(CSRC, s), "-o", o])
        objs.append(o)
    o = os.path.join(OUT, "cuda_emu.o")
    subprocess.check_call(["g++", "-std=c++20", "-O1", "-g", "-fPIC", "-I", HERE, "-c", os.path.join(HERE, "cuda_emu.cpp"), "-o", o])
    objs.append(o)
    subprocess.check_call(["g++", "-shared", "-o", LIB] + objs + ["-lpthread"])
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv))
