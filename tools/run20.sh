set -x
timeout 300 python tools/profile_c5.py 32 2>&1 | tail -2
cp voxvision_b200/lib/libvoxgpu.so /tmp/default.so; cp voxvision_b200/_variants/libvoxgpu_noearly.so voxvision_b200/lib/libvoxgpu.so
timeout 300 python tools/profile_c5.py 32 2>&1 | tail -2
cp /tmp/default.so voxvision_b200/lib/libvoxgpu.so
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_render_persistent -s 18 -c 1 -o gpurun_out/k_render_c5_r02 -f python tools/profile_c5.py 20 > gpurun_out/ncu_c5.log 2>&1; tail -2 gpurun_out/ncu_c5.log
