mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2a_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2a_pytest.log
tail -5 gpurun_out/r2a_pytest.log
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/r2a_bench_c2.json 2> gpurun_out/r2a_bench_c2.log; tail -c 2000 gpurun_out/r2a_bench_c2.json
for cfg in c2nr c4 c5; do timeout 600 python bench.py --config $cfg --no-cpu-baseline --e2e-steps 2 > gpurun_out/r2a_bench_$cfg.json 2> gpurun_out/r2a_bench_$cfg.log; tail -c 1500 gpurun_out/r2a_bench_$cfg.json; done
