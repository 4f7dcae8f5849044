set -x
timeout 500 python -m pytest tests -m gpu -x -q 2>&1 | tail -6
timeout 400 python tools/ab.py 1:0 1:16 1:24 1:32 1:48 2>&1 | grep "^AB" | tee gpurun_out/ab_feedback_r02g.log
timeout 200 python tools/trace_tiles.py 2>&1 | head -12 | cut -c1-250
