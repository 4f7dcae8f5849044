set -x
timeout 300 python -m pytest tests/test_gpu_round2.py -x -q 2>&1 | tail -4
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 8 --steps 3 > gpurun_out/bench_n8.log 2> gpurun_out/bench_n8.err; tail -c 600 gpurun_out/bench_n8.log; grep -i "error\|Traceback" gpurun_out/bench_n8.err | head -5
