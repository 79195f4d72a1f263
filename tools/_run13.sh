mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/r2n_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2n_pytest.log
B="python bench.py --no-cpu-baseline --no-e2e --no-extras --steps 10 --warmup 3"
: > gpurun_out/r2n_variants.txt
run() { echo "== $1" >> gpurun_out/r2n_variants.txt; env $2 timeout 300 $B --config ${3:-c2} 2>>gpurun_out/r2n_err.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['stage_ms'], d.get('parity'))" >> gpurun_out/r2n_variants.txt; }
run "c2 tensor-core" ""
run "c5 tensor-core" "" c5
run "c2q50 tensor-core" "" c2q50
run "c2q95 tensor-core" "" c2q95
cat gpurun_out/r2n_variants.txt; tail -5 gpurun_out/r2n_err.log
CMD="python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-extras"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:mcu_rgb -c 1 -f -o gpurun_out/r2n_tc_full $CMD > gpurun_out/r2n_ncu.log 2>&1; echo ncu rc=$?
