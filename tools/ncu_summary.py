"""Turn an `ncu --set full` report into the per-kernel table committed under profiles/ (development / evidence tool).

    python tools/ncu_summary.py gpurun_out/r2_final_full.ncu-rep > profiles/r2_ncu_summary.md
Reads the report with `ncu -i ... --page raw --csv` (works without a GPU).
"""
import csv
import io
import subprocess
import sys

COLS = [("gpu__time_duration.sum", "duration us", 1.0),
        ("dram__bytes_read.sum", "DRAM read MB", None),
        ("dram__bytes_write.sum", "DRAM write MB", None),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM % of peak", 1.0),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue active %", 1.0),
        ("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "ALU pipe %", 1.0),
        ("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "FMA pipe %", 1.0),
        ("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "LSU pipe %", 1.0),
        ("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "XU pipe %", 1.0),
        ("sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "FP64 pipe %", 1.0),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active %", 1.0),
        ("smsp__inst_executed.sum", "warp instr M", 1e-6),
        ("launch__registers_per_thread", "regs", 1.0),
        ("launch__grid_size", "grid", 1.0),
        ("launch__block_size", "block", 1.0),
        ("launch__shared_mem_per_block_dynamic", "dyn smem B", 1.0),
        ("launch__shared_mem_per_block_static", "static smem B", 1.0)]


def to_mb(v, unit):
    f = {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}.get(unit, None)
    return v * f if f else v


def main(path):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    print("| kernel | " + " | ".join(c[1] for c in COLS) + " |")
    print("|---|" + "---:|" * len(COLS))
    for r in rows[2:]:
        name = r[idx["Kernel Name"]].split("(")[0].replace("void ", "")
        cells = []
        for key, _, scale in COLS:
            if key not in idx or r[idx[key]] == "":
                cells.append("")
                continue
            v = float(r[idx[key]].replace(",", ""))
            v = to_mb(v, units[idx[key]]) if scale is None else v * scale
            if key == "gpu__time_duration.sum":
                v = v * {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(units[idx[key]], 1.0)
            cells.append(f"{v:.1f}" if abs(v) < 1e4 else f"{v:.0f}")
        print(f"| `{name}` | " + " | ".join(cells) + " |")


if __name__ == "__main__":
    main(sys.argv[1])
