mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-extras"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:mcu_rgb -c 1 -f -o gpurun_out/r2m_tc_full $CMD > gpurun_out/r2m_ncu.log 2>&1; echo rc=$?; tail -3 gpurun_out/r2m_ncu.log
