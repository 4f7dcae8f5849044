set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -15
for r in 0 4 8 16; do VOXGPU_RAYS_REFILL=$r python tools/profile_rays.py 1 4 2>&1 | tail -3; done
python bench.py --steps 3 > gpurun_out/bench_r02d.log 2> gpurun_out/bench_r02d.err; tail -c 6000 gpurun_out/bench_r02d.log; tail -5 gpurun_out/bench_r02d.err
python bench.py --impl reference --steps 2 --warmup 1 | cut -c1-400
compute-sanitizer --tool memcheck --error-exitcode 7 python __graft_entry__.py smoke > gpurun_out/sanitizer_memcheck.log 2>&1; echo memcheck rc $?; tail -4 gpurun_out/sanitizer_memcheck.log
compute-sanitizer --tool racecheck --error-exitcode 7 python __graft_entry__.py smoke > gpurun_out/sanitizer_racecheck.log 2>&1; echo racecheck rc $?; tail -4 gpurun_out/sanitizer_racecheck.log
