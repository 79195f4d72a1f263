mkdir -p gpurun_out
B="python bench.py --no-cpu-baseline --no-e2e --no-extras --steps 10 --warmup 3"
run() { echo "== $1" >> gpurun_out/r2f_variants.txt; env $2 timeout 600 $B --config ${3:-c2} 2>>gpurun_out/r2f_err.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['stage_ms'])" >> gpurun_out/r2f_variants.txt; }
run "sorted intervals (default)" ""
run "unsorted" "HJD_LIB_PATH=$PWD/tune/libhjd_nosort.so"
run "sorted q95" "" c2q95
run "unsorted q95" "HJD_LIB_PATH=$PWD/tune/libhjd_nosort.so" c2q95
run "sorted c5" "" c5
run "unsorted c5" "HJD_LIB_PATH=$PWD/tune/libhjd_nosort.so" c5
cat gpurun_out/r2f_variants.txt
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_round2.py tests/test_gpu_configs.py -m gpu -x -q 2>&1 | tail -3
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/r2f_bench_full.json 2> gpurun_out/r2f_bench_full.log; tail -c 2500 gpurun_out/r2f_bench_full.json; tail -3 gpurun_out/r2f_bench_full.log
