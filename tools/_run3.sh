mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2c_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2c_pytest.log
tail -25 gpurun_out/r2c_pytest.log
B="python bench.py --no-cpu-baseline --no-e2e --no-extras --steps 10 --warmup 3"
for cfg in c2 c2q50 c2q95 c4 c5 "c2nr --images 256"; do
  echo "== $cfg" >> gpurun_out/r2c_stage.txt
  timeout 600 $B --config $cfg 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['stage_ms'])" >> gpurun_out/r2c_stage.txt
done
cat gpurun_out/r2c_stage.txt
CMD="python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-extras"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'hjd_k_(entropy|mcu)' -c 2 -f -o gpurun_out/r2c_full $CMD > gpurun_out/r2c_ncu_full.log 2>&1
ls -la gpurun_out | tail -5
