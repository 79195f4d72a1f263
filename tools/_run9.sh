mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2i_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2i_pytest.log
tail -6 gpurun_out/r2i_pytest.log
timeout 900 python tools/_ss_tune.py > gpurun_out/r2i_ss_tune.txt 2>&1; cat gpurun_out/r2i_ss_tune.txt
timeout 600 python bench.py --config c2nr --no-cpu-baseline --no-extras --e2e-steps 3 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('c2nr', d['ms_per_step'], d['stage_ms'], d['e2e'])"
