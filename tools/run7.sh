set -x
timeout 400 python -m pytest tests -m gpu -x -q 2>&1 | tail -8
timeout 300 python tools/e2e_probe.py 2 4 6 8 > gpurun_out/e2e_probe_r02f.log 2>&1; cat gpurun_out/e2e_probe_r02f.log
VOXGPU_NO_BAND_FLAGS=1 timeout 200 python tools/e2e_probe.py 4 2>&1 | tail -2
timeout 300 python tools/ab.py --c5 1 2>&1 | grep "^AB"
