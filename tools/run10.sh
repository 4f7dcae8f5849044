set -x
timeout 500 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
( time timeout 900 python bench.py ) > gpurun_out/bench_n1.log 2> gpurun_out/bench_n1.err; tail -c 600 gpurun_out/bench_n1.log; tail -4 gpurun_out/bench_n1.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.log 2>&1; cut -c1-300 gpurun_out/bench_ref.log
VOX_BENCH_SKIP=cpu,configs,c5 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/launches_bench_r02.csv python bench.py --steps 1 --warmup 1 > gpurun_out/ncu_bench.log 2>&1; tail -2 gpurun_out/ncu_bench.log | cut -c1-200
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_render_persistent -s 5 -c 2 -o gpurun_out/k_render_r02 -f python tools/profile_render.py 0 0 0 0 0 0 90 90 > gpurun_out/ncu_render.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_rays -s 1 -c 1 -o gpurun_out/k_rays_stream_r02 -f python tools/profile_rays.py 1 2 > gpurun_out/ncu_rays2.log 2>&1
ls -la gpurun_out/*.ncu-rep
