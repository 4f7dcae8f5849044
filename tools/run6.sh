set -x
./tools/probes/waitvalue_probe 1 10000 0 2>&1 | tail -5; ./tools/probes/waitvalue_probe 2 10000 0 2>&1 | tail -3; ./tools/probes/waitvalue_probe 1 10000 1 2>&1 | tail -3
nvidia-smi topo -m | head -8
python -m pytest tests/test_gpu_round2.py -k "split_over_the_devices" -x -q 2>&1 | tail -8
python tools/peer_store_probe.py 120 > gpurun_out/peer_store_probe.log 2>&1; cat gpurun_out/peer_store_probe.log
ncu --query-metrics 2>/dev/null | grep -i "^nvl" | head -20 > gpurun_out/nvlink_metrics.txt; head -5 gpurun_out/nvlink_metrics.txt
ncu --metrics nvltx__bytes.sum,nvlrx__bytes.sum,gpu__time_duration.sum --clock-control none -k regex:k_render_persistent -c 12 --csv --log-file gpurun_out/peer_store_ncu.csv python tools/peer_store_probe.py 4 > gpurun_out/peer_store_ncu.log 2>&1; tail -5 gpurun_out/peer_store_ncu.log; grep -c nvl gpurun_out/peer_store_ncu.csv
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 > gpurun_out/bench_n2.log 2> gpurun_out/bench_n2.err; tail -c 5000 gpurun_out/bench_n2.log; tail -5 gpurun_out/bench_n2.err
