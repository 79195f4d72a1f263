mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/r2g_topo.txt 2>&1
nproc >> gpurun_out/r2g_topo.txt; free -g >> gpurun_out/r2g_topo.txt; numactl -H >> gpurun_out/r2g_topo.txt 2>&1; lscpu | grep -i "numa\|model name\|socket" >> gpurun_out/r2g_topo.txt
for n in 1 2 4 8; do timeout 300 python tools/d2h_probe.py --gpus $n >> gpurun_out/r2g_d2h_probe.jsonl 2>> gpurun_out/r2g_probe_err.log; done
timeout 300 python tools/d2h_probe.py --gpus 8 --no-numa >> gpurun_out/r2g_d2h_probe.jsonl 2>> gpurun_out/r2g_probe_err.log
cat gpurun_out/r2g_d2h_probe.jsonl | cut -c1-400
timeout 600 python -m pytest tests/test_gpu_round2.py -m gpu -x -q -k "multi or six or pipelined" 2>&1 | tail -3
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/r2g_bench_n8.json 2> gpurun_out/r2g_bench_n8.log
tail -c 3000 gpurun_out/r2g_bench_n8.json; tail -5 gpurun_out/r2g_bench_n8.log
