set -x
nvidia-smi -L | head -3
python -m pytest tests -m gpu -x -q 2>&1 | tail -15
python tools/kbench.py 0 8 16 24 32 48 2>&1 | tail -60 > gpurun_out/kbench_r02a.log
python tools/trace_tiles.py 0 24 2>&1 > gpurun_out/trace_r02a.log
tail -5 gpurun_out/kbench_r02a.log
