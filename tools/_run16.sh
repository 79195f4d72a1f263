mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-extras"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:mcu_rgb -c 1 -f -o gpurun_out/r2v_b_full $CMD > gpurun_out/r2v_ncu.log 2>&1; echo ncu rc=$?
HJD_BENCH_FLAGS=64 timeout 600 ncu --set full --clock-control none --import-source on -k regex:mcu_rgb -c 1 -f -o gpurun_out/r2v_o_full $CMD > gpurun_out/r2v_ncu2.log 2>&1; echo ncu rc=$?
