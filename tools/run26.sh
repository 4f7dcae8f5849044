for cost in 98304 40000; do
  echo "== COST=$cost"
  for i in 0 1 2 3 4 5 6 7; do
    CUDA_VISIBLE_DEVICES=$i VOXGPU_CLEAR_THREADS=2 PROBE_STEPS=24 VOXGPU_COPY_COST=$cost timeout 300 python tools/e2e_probe.py 0 > gpurun_out/probe_$i.log 2>&1 &
  done
  wait
  cat gpurun_out/probe_*.log | grep bands | sed 's/device.*BEST x6/BEST x6/'
  python - <<'P'
import re,glob
a=[0,0]; mn=[1e9,1e9]
for f in glob.glob('gpurun_out/probe_*.log'):
    m=re.findall(r'\((\d+) Mrays/s\)',open(f).read())
    for i in range(min(2,len(m))): a[i]+=int(m[i]); mn[i]=min(mn[i],int(m[i]))
print('SUM x1',a[0],'x6',a[1],'MIN x1',mn[0],'x6',mn[1])
P
done
