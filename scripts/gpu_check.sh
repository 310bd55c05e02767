#!/bin/bash
# One GPU-box visit: parity tests, smoke, bench (with the per-phase breakdown), optional walk statistics.
# Usage (from the repo root on the box):  bash scripts/gpu_check.sh [lowres] [ncu] [ncufull]
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" | tee -a gpurun_out/smoke.log
timeout 600 python bench.py --breakdown > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench rc=$?"
tail -c 1500 gpurun_out/bench_n1.json; grep "phase ms" gpurun_out/bench_n1.err
for a in "$@"; do
  if [ "$a" = "lowres" ]; then
    timeout 600 python bench.py --guidance lowres --no-cpu-baseline --breakdown > gpurun_out/bench_n1_lowres.json 2> gpurun_out/bench_n1_lowres.err
    grep "phase ms" gpurun_out/bench_n1_lowres.err; python -c "import json;d=json.load(open('gpurun_out/bench_n1_lowres.json'));print('lowres ms/step',d['ms_per_step'])"
  fi
  if [ "$a" = "ncu" ]; then
    timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv \
      python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
    python scripts/summarize_launches.py gpurun_out/launches.csv | head -30
  fi
done
tail -n 3 gpurun_out/pytest.log
for a in "$@"; do
  if [ "$a" = "ncufull" ]; then
    timeout 900 ncu --set full --clock-control none --import-source on --kernel-name regex:composite -c 2 -f -o gpurun_out/r01_full \
      python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1; echo "ncu full rc=$?"
  fi
done
