set -x
timeout 500 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
timeout 300 python tools/e2e_probe.py 2 4 6 2>&1 | tee gpurun_out/e2e_probe_r02i.log
VOXGPU_HOST_CLEAR=0 timeout 200 python tools/e2e_probe.py 4 2>&1 | tee -a gpurun_out/e2e_probe_r02i.log
VOXGPU_CLEAR_THREADS=1 timeout 200 python tools/e2e_probe.py 4 2>&1 | tee -a gpurun_out/e2e_probe_r02i.log
VOXGPU_CLEAR_THREADS=8 timeout 200 python tools/e2e_probe.py 4 2>&1 | tee -a gpurun_out/e2e_probe_r02i.log
