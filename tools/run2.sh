set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -15
python tools/kbench.py 0 2>&1 | tail -12 > gpurun_out/kbench_r02b.log
python tools/trace_tiles.py 0 2>&1 > gpurun_out/trace_r02b.log
python tools/configs_bench.py c3 c1 c4 modes > gpurun_out/configs_r02b.log 2>&1
tail -3 gpurun_out/kbench_r02b.log; cat gpurun_out/configs_r02b.log | cut -c1-900
