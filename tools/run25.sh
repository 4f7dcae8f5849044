timeout 600 python -m pytest tests/test_gpu_round2.py -x -q -k "visible or thread" 2>&1 | tail -2
for cost in 32768 98304 262144; do
  echo "== COST=$cost"
  VOXGPU_COPY_COST=$cost timeout 300 python tools/e2e_probe.py 0 2>&1 | tail -1
done
