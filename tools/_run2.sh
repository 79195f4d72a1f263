mkdir -p gpurun_out
B="python bench.py --no-cpu-baseline --no-e2e --no-extras --steps 10 --warmup 3"
for v in default cap4 w8; do
  for cfg in "c4" "c2nr --images 256"; do
    case $v in
      default) env= ;;
      cap4) env="HJD_SS_SYNC_PER_SM=4" ;;
      w8) env="HJD_LIB_PATH=$PWD/tune/libhjd_w8.so" ;;
    esac
    echo "== $v $cfg" >> gpurun_out/r2b_ss_variants.txt
    env $env timeout 600 $B --config $cfg 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['stage_ms'])" >> gpurun_out/r2b_ss_variants.txt
  done
done
cat gpurun_out/r2b_ss_variants.txt
C4="python bench.py --config c4 --no-cpu-baseline --no-e2e --no-extras --steps 2 --warmup 3"
NR="python bench.py --config c2nr --images 256 --no-cpu-baseline --no-e2e --no-extras --steps 2 --warmup 3"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2b_launches_c4.csv $C4 > gpurun_out/r2b_ncu_c4.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2b_launches_c2nr256.csv $NR > gpurun_out/r2b_ncu_nr.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'ss_|destuff' -c 8 -f -o gpurun_out/r2b_ss_full $NR > gpurun_out/r2b_ncu_ss_full.log 2>&1
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/r2b_bench_extras.json 2> gpurun_out/r2b_bench_extras.log; tail -c 3000 gpurun_out/r2b_bench_extras.json; tail -5 gpurun_out/r2b_bench_extras.log
