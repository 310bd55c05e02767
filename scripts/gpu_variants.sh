#!/bin/bash
# bench several builds of the library back to back (BDS_LIB selects the .so); prints the per-phase breakdown
mkdir -p gpurun_out
for lib in "$@"; do
  BDS_LIB=$lib timeout 600 python bench.py --breakdown --no-cpu-baseline --steps 10 > gpurun_out/bench_$lib.json 2> gpurun_out/bench_$lib.err
  echo "$lib rc=$? $(python -c "import json;d=json.load(open('gpurun_out/bench_$lib.json'));print(d['ms_per_step'])")"
  grep "phase ms" gpurun_out/bench_$lib.err
done
