mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2e_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2e_pytest.log
tail -15 gpurun_out/r2e_pytest.log
B="python bench.py --no-cpu-baseline --no-e2e --no-extras --steps 10 --warmup 3"
run() { echo "== $1" >> gpurun_out/r2e_variants.txt; env $2 timeout 600 $B --config ${3:-c2} 2>>gpurun_out/r2e_err.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['stage_ms'])" >> gpurun_out/r2e_variants.txt; }
run "2 variants (default)" ""
run "1 variant" "HJD_LIB_PATH=$PWD/tune/libhjd_v1.so"
run "2 variants q50" "" c2q50
run "1 variant q50" "HJD_LIB_PATH=$PWD/tune/libhjd_v1.so" c2q50
run "2 variants c5" "" c5
run "2 variants c4" "" c4
cat gpurun_out/r2e_variants.txt
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/r2e_bench_full.json 2> gpurun_out/r2e_bench_full.log; tail -c 1800 gpurun_out/r2e_bench_full.json; tail -3 gpurun_out/r2e_bench_full.log
