mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_round2.py -m gpu -x -q > gpurun_out/r2t_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2t_pytest.log
B="python bench.py --no-cpu-baseline --no-e2e --no-extras --steps 10 --warmup 3"
: > gpurun_out/r2t_variants.txt
run() { echo "== $1" >> gpurun_out/r2t_variants.txt; env $2 timeout 300 $B --config ${3:-c2} 2>>gpurun_out/r2t_err.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['stage_ms'], d.get('parity'))" >> gpurun_out/r2t_variants.txt; }
run "c2 tensor-core" ""
run "c2 cuda-core" "HJD_BENCH_FLAGS=64"
run "c5 tensor-core" "" c5
cat gpurun_out/r2t_variants.txt; tail -5 gpurun_out/r2t_err.log


CMD="python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-extras"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:mcu_rgb -c 1 -f -o gpurun_out/r2t_tc_full $CMD > gpurun_out/r2t_ncu.log 2>&1; echo ncu rc=$?
