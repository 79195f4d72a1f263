mkdir -p gpurun_out
tune/exp_idct_lanes > gpurun_out/r2j_exp_idct_lanes.json 2>&1; cat gpurun_out/r2j_exp_idct_lanes.json
tune/exp_idct_lanes 33554432 >> gpurun_out/r2j_exp_idct_lanes.json 2>&1; tail -1 gpurun_out/r2j_exp_idct_lanes.json
B="python bench.py --no-cpu-baseline --no-e2e --no-extras --steps 10 --warmup 3"
run() { echo "== $1" >> gpurun_out/r2j_variants.txt; env $2 timeout 600 $B --config ${3:-c2} 2>>gpurun_out/r2j_err.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['stage_ms'])" >> gpurun_out/r2j_variants.txt; }
run "c2nr maxr256" "" c2nr
run "c2nr maxr512" "HJD_LIB_PATH=$PWD/tune/libhjd_r512.so" c2nr
run "c4 maxr256" "" c4
run "c4 maxr512" "HJD_LIB_PATH=$PWD/tune/libhjd_r512.so" c4
cat gpurun_out/r2j_variants.txt
