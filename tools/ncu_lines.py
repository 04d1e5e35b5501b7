"""Aggregate an `ncu --page source --csv --print-source cuda,sass` export into per-source-line sample counts."""
import csv
import sys
from collections import defaultdict


def main(path, top=25):
    cur_file = None
    agg = defaultdict(lambda: [0, 0, ""])
    with open(path, newline="") as f:
        for row in csv.reader(f):
            if not row:
                continue
            if row[0] == "File Path":
                cur_file = row[1].split("/")[-1]
                continue
            if row[0] in ("Function Name", "Line No", "Kernel Name", "File Name"):
                continue
            if row[0].isdigit() and len(row) > 7:
                try:
                    samples = int(row[6])
                    inst = int(row[7])
                except ValueError:
                    continue
                k = (cur_file, int(row[0]))
                agg[k][0] += samples
                agg[k][1] += inst
                agg[k][2] = row[1].strip()[:110]
    tot = sum(v[0] for v in agg.values()) or 1
    toti = sum(v[1] for v in agg.values()) or 1
    print(f"total samples {tot}, total warp-instructions {toti}")
    for (fn, ln), (s, i, src) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
        print(f"{100 * s / tot:5.1f}% smp {100 * i / toti:5.1f}% inst  {fn}:{ln:<4d} {src}")


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 25)
