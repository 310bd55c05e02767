mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on --kernel-name regex:composite -c 2 -f -o gpurun_out/r01_full python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1; echo "ncu full rc=$?"; tail -3 gpurun_out/ncu_full.log
