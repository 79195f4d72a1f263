mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/r2_final_bench_n8.json 2> gpurun_out/r2_final_bench_n8.log; echo "rc=$?"; tail -c 600 gpurun_out/r2_final_bench_n8.json
timeout 300 python -m pytest tests/test_gpu_round2.py -m gpu -x -q -k "multi or device" 2>&1 | tail -3
