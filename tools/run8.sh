set -x
nvidia-smi topo -m | head -12
timeout 300 python -m pytest tests/test_gpu_round2.py -x -q 2>&1 | tail -5
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --steps 3 > gpurun_out/bench_n8.log 2> gpurun_out/bench_n8.err; tail -c 3000 gpurun_out/bench_n8.log; tail -5 gpurun_out/bench_n8.err
