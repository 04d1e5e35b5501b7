"""profiles/traffic.json from an `ncu --set full` report of tools/prof_kernels.py (development / evidence tool).

    python tools/traffic_from_ncu.py gpurun_out/r2_final_full.ncu-rep > profiles/traffic.json
DRAM bytes per frame of every detection stage = (dram__bytes_read.sum + dram__bytes_write.sum) / frames of the launch, summed over
the kernels of the stage (bench.py reads the file for roofline.traffic and roofline.kernels).
"""
import csv
import io
import json
import subprocess
import sys

STAGES = {"scan": (["scan_hot_vec32_kernel"], "hbm"),
          "group": (["form_clusters_kernel"], "latency (one CTA per frame)"),
          "filter": (["piece_filter_kernel"], "L1 data pipe / instruction issue: its DRAM rate is a description, not a roofline"),
          "borders": (["candidates_kernel", "borders_finalize_kernel"], "instruction issue / latency of dependent integer chains"),
          "finish": (["mark_active_kernel", "compact_tiles_kernel", "filter_tiles_kernel", "blobs_kernel"],
                     "latency (no frame of the workload needs the general path)"),
          "scan_tma": (["scan_tma_kernel"], "hbm")}
UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}


def main(path, frames=1024):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    seen = {}
    for r in rows[2:]:
        name = r[idx["Kernel Name"]].split("(")[0].replace("void ", "").split("<")[0]
        if name in seen:
            continue                                   # first launch of a kernel = the detection step of prof_kernels.py
        tot = 0.0
        for key in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
            tot += float(r[idx[key]].replace(",", "")) * UNIT[units[idx[key]]]
        seen[name] = tot
    out = {"_source": f"ncu --set full --clock-control none, tools/prof_kernels.py on the bench workload ({frames} frames of 2048x2048 per launch), "
                      f"{path}; dram_bytes_per_frame = (dram__bytes_read.sum + dram__bytes_write.sum) / {frames}, summed over the kernels of a stage"}
    for st, (ks, bound) in STAGES.items():
        have = [k for k in ks if k in seen]
        if have:
            out[st] = {"kernels": have, "dram_bytes_per_frame": sum(seen[k] for k in have) / frames, "frames": frames, "bound": bound}
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main(sys.argv[1])
