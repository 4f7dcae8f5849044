for c in 5 0; do
  echo "== BUSY_CTAS=$c"
  VOXGPU_K1_BUSY_CTAS=$c VOX_BENCH_SKIP=cpu,configs timeout 400 python bench.py > gpurun_out/b_busy$c.log 2> gpurun_out/b_busy$c.err; tail -2 gpurun_out/b_busy$c.err
  VOXGPU_K1_BUSY_CTAS=$c AB_TAG=busy$c timeout 300 python tools/ab.py 2>&1 | grep "^AB" | cut -c1-300
done
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -2
