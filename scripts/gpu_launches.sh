#!/bin/bash
# ncu launch list (gpu__time_duration) of a short bench run: per-launch times are cold-cache and serialised - compare shares.
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches.csv \
  python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
echo "ncu rc=$?"
python scripts/summarize_launches.py gpurun_out/launches.csv | head -40
