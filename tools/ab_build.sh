#!/bin/bash
# Build-time variants of libvoxgpu.so into voxvision_b200/_variants/ (they travel to the GPU box with the snapshot):
#   tools/ab_build.sh name "-DVOXGPU_X=1 ..." [name "flags" ...]
cd "$(dirname "$0")/../voxvision_b200"
mkdir -p _variants
while [ $# -ge 2 ]; do
  name=$1; flags=$2; shift 2
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -fmad=false -Xcompiler -fPIC,-fvisibility=hidden,-fopenmp \
       -I../include -Icsrc/gpu $flags -shared -o _variants/libvoxgpu_$name.so csrc/gpu/voxgpu.cu -lgomp || exit 1
  echo "built $name ($flags): $(cuobjdump -res-usage _variants/libvoxgpu_$name.so 2>/dev/null | grep -A1 'k_render_persistentILb0ELj64ELb0' | grep -o 'REG:[0-9]*\|STACK:[0-9]*\|SHARED:[0-9]*' | tr '\n' ' ')"
done
