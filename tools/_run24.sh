mkdir -p gpurun_out
P=${1:-r3g}
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_round2.py tests/test_gpu_configs.py -m gpu -x -q > gpurun_out/${P}_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/${P}_pytest.log
B="python bench.py --no-cpu-baseline --no-e2e --no-extras --steps 10 --warmup 3"
: > gpurun_out/${P}_variants.txt
run() { echo "== $1" >> gpurun_out/${P}_variants.txt; env $2 timeout 300 $B $3 2>>gpurun_out/${P}_err.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['stage_ms'])" >> gpurun_out/${P}_variants.txt; }
for n in 128 256 512 1024; do run "c2 images=$n" "" "--config c2 --images $n"; done
cat gpurun_out/${P}_variants.txt; tail -3 gpurun_out/${P}_err.log
