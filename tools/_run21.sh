mkdir -p gpurun_out
P=${1:-r3c}
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${P}_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/${P}_pytest.log
timeout 600 python bench.py > gpurun_out/${P}_bench.json 2> gpurun_out/${P}_bench.log; echo "bench rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/${P}_bench.json'))
print(d['value'], d['ms_per_step'], d['stage_ms'], d['roofline'], d['e2e']['value'], d['parity'], d['c1'], d['file_to_bmp']['images_per_s'], d['c4']['ms_per_step'], d['c5']['ms_per_step'], d['strong']['ms_per_step'])"
