mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_round2.py -m gpu -x -q > gpurun_out/r2q_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2q_pytest.log
B="python bench.py --no-cpu-baseline --no-e2e --no-extras --steps 10 --warmup 3"
: > gpurun_out/r2q_variants.txt
run() { echo "== $1" >> gpurun_out/r2q_variants.txt; env $2 timeout 300 $B --config ${3:-c2} 2>>gpurun_out/r2q_err.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['stage_ms'], d.get('parity'))" >> gpurun_out/r2q_variants.txt; }
run "c2 tensor-core" ""
run "c2 cuda-core" "HJD_BENCH_FLAGS=64"
run "c5 tensor-core" "" c5
cat gpurun_out/r2q_variants.txt; tail -5 gpurun_out/r2q_err.log


