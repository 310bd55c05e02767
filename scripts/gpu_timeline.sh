#!/bin/bash
# kernel timeline of one bench step (torch.profiler / CUPTI): long kernels and gaps
# usage: gpu_timeline.sh [library.so] [extra bench.py arguments]
mkdir -p gpurun_out
BDS_LIB=${1:-libbds_b200.so} BDS_TIMELINE=gpurun_out/timeline.json python bench.py --steps 3 --no-cpu-baseline ${@:2} > /dev/null 2> gpurun_out/timeline.err
python scripts/timeline_summary.py gpurun_out/timeline.json
gzip -f gpurun_out/timeline.json
