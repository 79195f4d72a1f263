mkdir -p gpurun_out
P=${1:-r2z}
B="python bench.py --no-cpu-baseline --no-e2e --no-extras --steps 10 --warmup 3"
: > gpurun_out/${P}_variants.txt
run() { echo "== $1" >> gpurun_out/${P}_variants.txt; env $2 timeout 300 $B --config ${3:-c2} 2>>gpurun_out/${P}_err.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['stage_ms'], d.get('parity'))" >> gpurun_out/${P}_variants.txt; }
for v in $VARIANTS; do
  run "c2 tc $v" "HJD_BENCH_FLAGS=128 HJD_LIB_PATH=$PWD/tune/libhjd_$v.so"
done
run "c2 tc default" "HJD_BENCH_FLAGS=128"
cat gpurun_out/${P}_variants.txt; tail -5 gpurun_out/${P}_err.log
