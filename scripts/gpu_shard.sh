#!/bin/bash
mkdir -p gpurun_out
python scripts/shard_bench.py 8 32768 30 2>&1 | tail -2
ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 40 --csv --log-file gpurun_out/shard_launches.csv python scripts/shard_bench.py 8 32768 8 > /dev/null 2>&1
python - <<'PY'
import csv
rows=list(csv.reader(open('gpurun_out/shard_launches.csv')))
i=[k for k,r in enumerate(rows) if r and r[0]=="ID"][0]
hdr=rows[i]; kn=hdr.index("Kernel Name"); mv=hdr.index("Metric Value"); gs=hdr.index("Grid Size")
for r in rows[i+1:i+26]: print(f"{float(r[mv].replace(',',''))/1e3:9.1f} us  {r[gs]:>16s}  {r[kn][:90]}")
PY
