mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2h_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2h_pytest.log
tail -6 gpurun_out/r2h_pytest.log
B="python bench.py --no-cpu-baseline --no-e2e --no-extras --steps 10 --warmup 3"
run() { echo "== $1" >> gpurun_out/r2h_variants.txt; env $2 timeout 600 $B --config ${3:-c2} 2>>gpurun_out/r2h_err.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['stage_ms'])" >> gpurun_out/r2h_variants.txt; }
run "packed epilogue + once-per-block bad-code check (default)" ""
run "scalar epilogue, same entropy kernel" "HJD_LIB_PATH=$PWD/tune/libhjd_nopair.so"
run "default q50" "" c2q50
run "default q95" "" c2q95
run "default c4" "" c4
run "default c5" "" c5
run "default c2nr" "" c2nr
cat gpurun_out/r2h_variants.txt
