mkdir -p gpurun_out
bash tools/refresh_profiles.sh all r2_final 2> gpurun_out/r2_final_refresh.log
CMD="python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-extras"
HJD_BENCH_FLAGS=64 timeout 600 ncu --set full --clock-control none --import-source on -k regex:mcu_rgb -c 1 -f -o gpurun_out/r2_final_cuda_core_full $CMD > gpurun_out/r2_final_ncu_cc.log 2>&1
timeout 900 python tools/soak_parity.py 4000 7 > gpurun_out/r2_final_soak.txt 2>&1; tail -2 gpurun_out/r2_final_soak.txt
timeout 900 python tools/soak_parity.py 60 8 2400 1800 >> gpurun_out/r2_final_soak.txt 2>&1; tail -1 gpurun_out/r2_final_soak.txt
