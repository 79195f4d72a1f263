mkdir -p gpurun_out
P=${1:-r2w}
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_round2.py tests/test_gpu_configs.py -m gpu -x -q > gpurun_out/${P}_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/${P}_pytest.log
B="python bench.py --no-cpu-baseline --no-e2e --no-extras --steps 10 --warmup 3"
: > gpurun_out/${P}_variants.txt
run() { echo "== $1" >> gpurun_out/${P}_variants.txt; env $2 timeout 300 $B --config ${3:-c2} 2>>gpurun_out/${P}_err.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['stage_ms'], d.get('parity'))" >> gpurun_out/${P}_variants.txt; }
run "c2 cuda-core" ""
run "c2 tensor-core" "HJD_BENCH_FLAGS=128"
run "c5 cuda-core" "" c5
run "c5 tensor-core" "HJD_BENCH_FLAGS=128" c5
cat gpurun_out/${P}_variants.txt; tail -5 gpurun_out/${P}_err.log
CMD="python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-extras"
HJD_BENCH_FLAGS=128 timeout 600 ncu --set full --clock-control none --import-source on -k regex:mcu_rgb -c 1 -f -o gpurun_out/${P}_tc_full $CMD > gpurun_out/${P}_ncu.log 2>&1; echo ncu rc=$?
