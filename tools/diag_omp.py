import os, sys, time, subprocess
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
CASES = {
 "port_t1": ("port", 1, False), "port_t16": ("port", 16, False), "ref_t1": ("ref", 1, False),
 "ref_t16": ("ref", 16, False), "ref_t16_cuda": ("ref", 16, True), "ref_t4_cuda": ("ref", 4, True),
 "port_t16_cuda": ("port", 16, True),
}
if len(sys.argv) > 1:
    which, th, cuda = CASES[sys.argv[1]]
    import numpy as np
    from oracle import orc
    from tools import datasets as ds
    import voxvision_b200 as vv
    P = orc.Oracle(which)
    vox = ds.volume_voxels("skull")
    if cuda:
        ctx = vv.Context(1920, 1080)
        T = vv.Tree.build(vox)
    TP = P.tree(vox)
    pos, look = ds.orbit_pose(0)
    ocam = ds.setup_camera(P.camera(0), pos, look, 1.2, 1920, 1080)
    t = time.time(); r = TP.render(ocam, 1920, 1080, 0, threads=th); print(sys.argv[1], "ok %.3fs" % (time.time() - t), (r[1] >= 0).sum(), flush=True)
    t = time.time(); r = TP.render(ocam, 1920, 1080, 0, threads=th); print(sys.argv[1], "2nd %.3fs" % (time.time() - t), flush=True)
else:
    os.system("env | grep -i -E 'omp|gomp|kmp' ; ulimit -a | head -20; cat /sys/fs/cgroup/cpu.max 2>/dev/null; nproc; python -c 'import os; print(os.sched_getaffinity(0))'")
    for c in CASES:
        t = time.time()
        try:
            p = subprocess.run([sys.executable, "-u", __file__, c], timeout=45, capture_output=True, text=True)
            print(c, "rc", p.returncode, "%.1fs" % (time.time() - t), p.stdout.strip().replace("\n", " | "), p.stderr[-300:], flush=True)
        except subprocess.TimeoutExpired as e:
            print(c, "TIMEOUT", (e.stdout or b"")[-200:], flush=True)
