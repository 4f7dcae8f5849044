set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -15
python tools/kbench.py 2>&1 | tail -12 > gpurun_out/kbench_r02c.log
tail -3 gpurun_out/kbench_r02c.log
python tools/trace_tiles.py 2>&1 > gpurun_out/trace_r02c.log
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_render_r02c.csv python tools/profile_render.py 0 0 0 0 90 90 90 90 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_rays -s 1 -c 1 -o gpurun_out/k_rays_asgiven_r02c -f python tools/profile_rays.py 0 2 > gpurun_out/ncu_rays0.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_rays -s 1 -c 1 -o gpurun_out/k_rays_sorted_r02c -f python tools/profile_rays.py 1 2 > gpurun_out/ncu_rays1.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/launches_rays_r02c.csv python tools/profile_rays.py 1 3 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_render_persistent -s 5 -c 2 -o gpurun_out/k_render_r02c -f python tools/profile_render.py 0 0 0 0 0 0 90 90 > gpurun_out/ncu_render.log 2>&1
python tools/configs_bench.py modes > gpurun_out/modes_r02c.log 2>&1
ls -la gpurun_out/*.ncu-rep
