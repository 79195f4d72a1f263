mkdir -p gpurun_out
P=${1:-r3a}
B="python bench.py --no-cpu-baseline --no-e2e --no-extras --steps 10 --warmup 3"
: > gpurun_out/${P}_variants.txt
run() { echo "== $1" >> gpurun_out/${P}_variants.txt; env $2 timeout 300 $B --config ${3:-c2} 2>>gpurun_out/${P}_err.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['stage_ms'], d.get('parity'))" >> gpurun_out/${P}_variants.txt; }
for c in c4 c2q50 c2q95; do
  run "$c cuda-core" "" $c
  run "$c tensor-core" "HJD_BENCH_FLAGS=128" $c
done
cat gpurun_out/${P}_variants.txt; tail -5 gpurun_out/${P}_err.log
