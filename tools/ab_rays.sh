#!/bin/bash
# A/B of build-time variants on the batch configs (C3 rays, C4 balls): tools/configs_bench.py per variant.
cd "$(dirname "$0")/.."
L=voxvision_b200/lib/libvoxgpu.so
cp $L /tmp/libvoxgpu_default.so
for so in voxvision_b200/_variants/libvoxgpu_*.so; do
  name=$(basename $so .so); name=${name#libvoxgpu_}
  cp $so $L
  python tools/configs_bench.py c3 > /tmp/ab_c.log 2>&1
  echo "$name | $(python -c "
import json
for l in open('/tmp/ab_c.log'):
    if l.startswith('{'):
        d=json.loads(l); print(d['config'][:2], 'device', d['device'], 'kernel_ms', d['kernel_ms'], 'e2e', d['e2e_host_buffers'], end=' ; ')
" 2>&1 | tail -1)"
done
cp /tmp/libvoxgpu_default.so $L
