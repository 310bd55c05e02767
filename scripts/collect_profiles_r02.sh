#!/bin/bash
# Copy / derive the round-2 evidence from gpurun_out/ (written by scripts/gpu_profiles_r02.sh and the multi-GPU runs)
# into profiles/.  Needs ncu, cuobjdump, nvdisasm on PATH; no GPU.
set -e
cd "$(dirname "$0")/.."
for f in r02_bench_n1 r02_bench_reference_arm r02_bench_cfg1_500k_1cam r02_bench_n1_lowres_guidance r02_bench_masked_rerender \
         r02_bench_n2 r02_bench_n4 r02_bench_n8 r02_bench_n8_8M r02_bench_n8_dense; do
  [ -s gpurun_out/$f.json ] && cp gpurun_out/$f.json profiles/$f.json
done
if [ -s gpurun_out/r02_launches.csv ]; then
  cp gpurun_out/r02_launches.csv profiles/r02_launches_bench_step.csv
  python scripts/summarize_launches.py gpurun_out/r02_launches.csv > profiles/r02_launches_summary.txt
fi
if [ -s gpurun_out/r02_full.ncu-rep ]; then
  python scripts/ncu_key_metrics.py gpurun_out/r02_full.ncu-rep > profiles/r02_ncu_full_key_metrics.txt
  python scripts/ncu_traffic.py gpurun_out/r02_full.ncu-rep "profiles/r02_ncu_full_key_metrics.txt (ncu --set full, scripts/gpu_profiles_r02.sh)" > /dev/null
  { python scripts/ncu_source_lines.py gpurun_out/r02_full.ncu-rep composite_bwd composite 40; python scripts/ncu_source_lines.py gpurun_out/r02_full.ncu-rep composite_fwd composite 30; } > profiles/r02_ncu_source_lines_composite.txt
fi
python scripts/sass_markers.py > profiles/r02_sass_markers.txt
ls -la profiles | grep r02
