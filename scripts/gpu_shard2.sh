#!/bin/bash
for bc in 1 2 4 6 8 12 16; do echo -n "BWD_CHUNKS=$bc  "; VPA_BWD_CHUNKS=$bc python scripts/shard_bench.py 8 32768 20 2>&1 | tail -1; done
for fc in 2 4 6 9 14 18; do echo -n "FWD1_CHUNKS=$fc  "; VPA_FWD1_CHUNKS=$fc python scripts/shard_bench.py 8 32768 20 2>&1 | tail -1; done
