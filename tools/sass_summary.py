"""Opcode summary of libmocap_b200.so per kernel (evidence tool): what the claims about loads / TMA / barriers compile to.

    python tools/sass_summary.py > profiles/r2_sass_opcodes.md        (cuobjdump -sass, no GPU needed)
"""
import collections
import os
import re
import subprocess
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(REPO, "mocapv2_b200", "libmocap_b200.so")
WATCH = [("UTMALDG", r"^UTMALDG"), ("SYNCS (mbarrier)", r"^SYNCS"), ("UBLKCP", r"^UBLKCP"), ("LDGSTS (cp.async)", r"^LDGSTS"),
         ("LDG.*256", r"^LDG\S*\.256"), ("LDG.*128", r"^LDG\S*\.128"), ("LDG other", r"^LDG(?!STS)(?!\S*\.(128|256))"), ("LDS.128", r"^LDS\S*\.128"), ("LDS other", r"^LDS(?!\S*\.128)"),
         ("STS", r"^STS"), ("REDUX", r"^REDUX"), ("VOTE", r"^VOTE"), ("SHFL", r"^SHFL"), ("ATOMS/ATOMG/RED", r"^(ATOMS|ATOMG|ATOM|RED)\b|^(ATOMS|ATOMG|RED)\."),
         ("IDP", r"^IDP"), ("VIMNMX", r"^VIMNMX"), ("POPC/FLO", r"^(POPC|FLO|BREV)"), ("MUFU", r"^MUFU"), ("FFMA", r"^FFMA"), ("FMUL/FADD", r"^(FMUL|FADD)"),
         ("DFMA/DMUL/DADD", r"^(DFMA|DMUL|DADD)"), ("LOP3", r"^LOP3"), ("IMAD*", r"^IMAD"), ("BAR", r"^BAR"), ("HMMA/UTC*MMA (tensor)", r"^(HMMA|UTC\w*MMA|IMMA)")]


def main():
    out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    kernels = collections.OrderedDict()
    cur = None
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            kernels[cur] = collections.Counter()
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4}\*/\s+(.*?);", line)
        if m and cur:
            ins = re.sub(r"^@!?U?P\w+\s+", "", m.group(1).strip())
            op = ins.split()[0] if ins else ""
            kernels[cur]["total"] += 1
            for name, pat in WATCH:
                if re.search(pat, op):
                    kernels[cur][name] += 1
    demangle = subprocess.run(["cu++filt"] + list(kernels), capture_output=True, text=True).stdout.splitlines()
    names = [d.split("(")[0].replace("void ", "") for d in demangle] if len(demangle) == len(kernels) else list(kernels)
    print("Static SASS opcode counts per kernel of mocapv2_b200/libmocap_b200.so (`cuobjdump -sass`, sm_100a).  Tensor-core opcodes are absent by design:")
    print("the path is byte / bit work and tiny independent solves (DESIGN.md section 4).\n")
    print("| kernel | total | " + " | ".join(n for n, _ in WATCH) + " |")
    print("|---|---:|" + "---:|" * len(WATCH))
    for (k, c), nm in zip(kernels.items(), names):
        print(f"| `{nm}` | {c['total']} | " + " | ".join(str(c[n]) if c[n] else "" for n, _ in WATCH) + " |")


if __name__ == "__main__":
    main()
