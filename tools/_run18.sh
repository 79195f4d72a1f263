mkdir -p gpurun_out
P=${1:-r2x}
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_round2.py -m gpu -x -q > gpurun_out/${P}_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/${P}_pytest.log
B="python bench.py --no-cpu-baseline --no-e2e --no-extras --steps 10 --warmup 3"
: > gpurun_out/${P}_variants.txt
run() { echo "== $1" >> gpurun_out/${P}_variants.txt; env $2 timeout 300 $B --config ${3:-c2} 2>>gpurun_out/${P}_err.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['stage_ms'], d.get('parity'))" >> gpurun_out/${P}_variants.txt; }
run "c2" ""
run "c5" "" c5
run "c2q50" "" c2q50
run "c2q95" "" c2q95
run "c4" "" c4
cat gpurun_out/${P}_variants.txt; tail -5 gpurun_out/${P}_err.log
