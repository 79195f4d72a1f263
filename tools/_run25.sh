mkdir -p gpurun_out
P=${1:-r2_final2}
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${P}_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/${P}_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${P}_smoke.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/${P}_smoke.log
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${P}_bench_reference.json 2> gpurun_out/${P}_bench_reference.log; echo "ref rc=$?"
timeout 600 python bench.py > gpurun_out/${P}_bench.json 2> gpurun_out/${P}_bench.log; echo "bench rc=$?"
python -c "
import json; d=json.load(open('gpurun_out/${P}_bench.json'))
print(d['value'], d['ms_per_step'], d['stage_ms'], d['roofline']['frac'], d['roofline']['whole_step_frac'], d['e2e']['value'], d['parity']['mismatches'], d['config']['idct_kernel'][:12], d['strong']['ms_per_step'], d['c4']['ms_per_step'], d['c5']['ms_per_step'], d['file_to_bmp']['images_per_s'])"
