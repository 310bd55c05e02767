#!/bin/bash
# Round-2 evidence run on one B200 (gpurun): tests, smoke, bench lines of the single-GPU configs, ncu launch list, ncu
# full capture of the two composite kernels.  Everything lands in gpurun_out/; scripts/collect_profiles_r02.sh copies
# the summaries into profiles/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?"; tail -n 2 gpurun_out/pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"
timeout 600 python bench.py --breakdown > gpurun_out/r02_bench_n1.json 2> gpurun_out/r02_bench_n1.err; echo "bench rc=$?"; grep "phase ms" gpurun_out/r02_bench_n1.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02_bench_reference_arm.json 2> gpurun_out/r02_bench_reference_arm.err; echo "reference arm rc=$?"
timeout 300 python bench.py --n-gauss 500000 --cams 1 --no-cpu-baseline --breakdown --steps 20 > gpurun_out/r02_bench_cfg1_500k_1cam.json 2> gpurun_out/r02_bench_cfg1.err; grep "phase ms" gpurun_out/r02_bench_cfg1.err
timeout 300 python bench.py --guidance lowres --no-cpu-baseline --breakdown > gpurun_out/r02_bench_n1_lowres_guidance.json 2> gpurun_out/r02_bench_lowres.err; grep "phase ms" gpurun_out/r02_bench_lowres.err
timeout 300 python scripts/bench_masked.py > gpurun_out/r02_bench_masked_rerender.json 2> gpurun_out/masked.err; tail -n 1 gpurun_out/r02_bench_masked_rerender.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02_launches.csv \
  python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1; echo "ncu list rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on --kernel-name regex:composite -c 2 -f -o gpurun_out/r02_full \
  python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1; echo "ncu full rc=$?"
