mkdir -p gpurun_out
B="python bench.py --no-cpu-baseline --no-e2e --no-extras --steps 10 --warmup 3"
run() { echo "== $1" >> gpurun_out/r2d_variants.txt; env $2 timeout 600 $B --config ${3:-c2} 2>>gpurun_out/r2d_err.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['stage_ms'])" >> gpurun_out/r2d_variants.txt; }
run "v3 L2=32" ""
run "v3 L2=64" "HJD_L2_FETCH=64"
run "v3 L2=128" "HJD_L2_FETCH=128"
run "v2 L2=32" "HJD_LIB_PATH=$PWD/tune/libhjd_v2.so"
run "v1 L2=32" "HJD_LIB_PATH=$PWD/tune/libhjd_v1.so"
run "noblast (timing only)" "HJD_LIB_PATH=$PWD/tune/libhjd_noblast.so"
run "v3 q50" "" c2q50
run "v3 q95" "" c2q95
run "v3 c5" "" c5
cat gpurun_out/r2d_variants.txt
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_round2.py -m gpu -x -q 2>&1 | tail -3
CMD="python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-extras"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'hjd_k_(entropy|mcu)' -c 2 -f -o gpurun_out/r2d_full $CMD > gpurun_out/r2d_ncu_full.log 2>&1
