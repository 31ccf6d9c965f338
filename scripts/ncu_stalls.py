"""Summarise an .ncu-rep here (no GPU): headline metrics + top stall locations per kernel.
    python scripts/ncu_stalls.py gpurun_out/prof_sweep.ncu-rep [min_pct]"""
import csv, subprocess, sys, io
rep = sys.argv[1]; minpct = float(sys.argv[2]) if len(sys.argv) > 2 else 0.5
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr = rows[0]
want = ["gpu__time_duration.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "sm__cycles_elapsed.avg", "lts__t_bytes.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smsp__inst_executed_pipe_xu.sum",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "lts__t_sectors_op_read.sum", "lts__t_sectors.sum", "sm__cycles_elapsed.avg.per_second",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "l1tex__m_xbar2l1tex_read_bytes.sum", "launch__registers_per_thread", "lts__t_sector_hit_rate.pct"]
for i, h in enumerate(hdr):
    if h in want or h == "Kernel Name":
        print(f"{h:75s}", [r[i][:40] for r in rows[1:]])
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
kern = []; cur = None
for r in rows:
    if r and r[0] == "Kernel Name":
        cur = {"name": r[1], "rows": []}; kern.append(cur); continue
    if cur is not None: cur["rows"].append(r)
for k in kern:
    hdr = k["rows"][0]; data = k["rows"][1:]
    i_src = hdr.index("Source"); i_s = hdr.index("# Samples"); i_ex = hdr.index("Instructions Executed")
    tot = sum(int(d[i_s] or 0) for d in data)
    sc = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
    print("=====", k["name"][:70], "samples", tot)
    for n, d in enumerate(data):
        sm = int(d[i_s] or 0)
        if sm / max(tot, 1) * 100 >= minpct:
            st = sorted(((int(d[i] or 0), hdr[i][6:]) for i in sc), reverse=True)[:2]
            print(f"{n:4d} {sm/tot*100:5.2f}% ex={d[i_ex]:>9} {d[i_src][:80]:80s} {st}")
