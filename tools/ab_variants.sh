#!/bin/bash
# A/B harness for build-time kernel variants: for each voxvision_b200/_variants/libvoxgpu_<name>.so,
# swap it in, check the golden C2 frames (tools/kbench.py) and run bench.py.  Restores the default library.
cd "$(dirname "$0")/.."
L=voxvision_b200/lib/libvoxgpu.so
cp $L /tmp/libvoxgpu_default.so
for so in voxvision_b200/_variants/libvoxgpu_*.so; do
  name=$(basename $so .so); name=${name#libvoxgpu_}
  cp $so $L
  python tools/kbench.py 1 > /tmp/ab_kb.log 2>&1
  kb=$(tail -1 /tmp/ab_kb.log | cut -c1-90)
  python bench.py --steps 2 > /tmp/ab_bench.log 2>&1
  bv=$(python -c "
import json
d=json.loads(open('/tmp/ab_bench.log').read().strip().splitlines()[-1]); print('value', d['value'], 'e2e', d['e2e']['value'], 'kernel_ms', d['roofline']['kernel_ms_mean'])
" 2>&1 | tail -1)
  par=$(grep -c "parity OK" /tmp/ab_kb.log)
  echo "$name | parityOK $par/8 | $kb | $bv"
done
cp /tmp/libvoxgpu_default.so $L
