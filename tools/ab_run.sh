#!/bin/bash
# For each voxvision_b200/_variants/libvoxgpu_<name>.so: swap it in and run tools/ab.py "$@".  Restores the default library.
cd "$(dirname "$0")/.."
L=voxvision_b200/lib/libvoxgpu.so
cp $L /tmp/libvoxgpu_default.so
for so in voxvision_b200/_variants/libvoxgpu_*.so; do
  name=$(basename $so .so); name=${name#libvoxgpu_}
  cp $so $L
  AB_TAG=$name python tools/ab.py "$@" 2>&1 | grep "^AB\|Error\|error" 
done
cp /tmp/libvoxgpu_default.so $L
