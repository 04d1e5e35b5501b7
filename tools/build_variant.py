"""Build a variant of libmocap_b200.so with extra -D flags into mocapv2_b200/_variants/<name>.so (development A/B tool).

    python tools/build_variant.py <name> [-DFLAG=VALUE ...]
"""
import os
import subprocess
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
from mocapv2_b200 import build as B  # noqa: E402


def main():
    name, flags = sys.argv[1], sys.argv[2:]
    out_dir = os.path.join(B.HERE, "_variants")
    os.makedirs(out_dir, exist_ok=True)
    lib = os.path.join(out_dir, name + ".so")
    cmd = [os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")] + B.NVCC_FLAGS + flags + ["-o", lib] + [os.path.join(B.CSRC, s) for s in B.SOURCES]
    out = subprocess.run(cmd, capture_output=True, text=True)
    open(lib + ".ptxas.log", "w").write(out.stdout + out.stderr)
    if out.returncode != 0:
        sys.stderr.write(out.stdout + out.stderr)
        raise SystemExit(1)
    for line in (out.stdout + out.stderr).splitlines():
        if "piece_filter" in line or "borders_finalize" in line:
            print(line.strip()[:160])
    print(lib)


if __name__ == "__main__":
    main()
