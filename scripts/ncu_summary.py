"""Summarise `ncu --set full` reports as a markdown table (one column per captured launch).
    python scripts/ncu_summary.py gpurun_out/prof_pair.ncu-rep [more.ncu-rep ...] > profiles/rNN_ncu_full_xx.md
Numbers under a profiler are for attribution only; bench values come from bench.py."""
import csv
import subprocess
import sys

METRICS = [
    ("duration", "gpu__time_duration.sum"),
    ("SM clock during capture", "sm__cycles_elapsed.avg.per_second"),
    ("tensor pipe active (% of peak sustained active)", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"),
    ("XU (MUFU) pipe instr (% of peak)", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active"),
    ("issue slots active", "sm__inst_issued.avg.pct_of_peak_sustained_active"),
    ("dram read", "dram__bytes_read.sum"),
    ("dram write", "dram__bytes_write.sum"),
    ("dram read rate", "dram__bytes_read.sum.per_second"),
    ("dram write rate", "dram__bytes_write.sum.per_second"),
    ("L2 sectors", "lts__t_sectors.sum"),
    ("shared-memory bank conflicts (LSU)", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"),
    ("registers/thread", "launch__registers_per_thread"),
    ("grid", "launch__grid_size"),
    ("block", "launch__block_size"),
    ("dyn smem", "launch__shared_mem_per_block_dynamic"),
    ("warps active", "sm__warps_active.avg.pct_of_peak_sustained_active"),
]


def load(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    start = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
    return rows[start], rows[start + 1], rows[start + 2:]


def main():
    cols = []
    for path in sys.argv[1:]:
        hdr, units, rows = load(path)
        kn = hdr.index("Kernel Name")
        for r in rows:
            name = r[kn].split("(")[0].replace("void ", "")
            vals = {}
            for label, m in METRICS:
                if m in hdr:
                    i = hdr.index(m)
                    v = r[i]
                    try:
                        v = f"{float(v.replace(',', '')):.4g}"
                    except ValueError:
                        pass
                    vals[label] = f"{v} {units[i]}".strip()
                else:
                    vals[label] = "n/a"
            cols.append((name, vals))
    print("| metric | " + " | ".join(f"`{n}`" for n, _ in cols) + " |")
    print("|---|" + "---|" * len(cols))
    for label, _ in METRICS:
        print(f"| {label} | " + " | ".join(v[label] for _, v in cols) + " |")


if __name__ == "__main__":
    main()
