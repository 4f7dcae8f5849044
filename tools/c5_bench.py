#!/usr/bin/env python
"""C5 (BASELINE.json configs[4]): a synthetic 1024^3 sparse volume (~100 M voxels,
union of balls with hash holes) rendered 7680x4320, the screen split across 1/2/4/8 B200.

    python tools/c5_bench.py [--frames 16] [--steps 3] [--check]                       one GPU
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
           --master-port P tools/c5_bench.py ...                                      N GPUs

One process per GPU.  The tree is built ONCE, on rank 0 (vox_make_tree, host), and replicated:
its flat arrays are broadcast over NCCL and imported on every rank (multi.broadcast_flat_tree).
Each frame is cut into world*BANDS_PER_RANK horizontal bands dealt round-robin (hit density is
uneven); a rank renders its bands with voxgpu_render straight into ONE frame buffer that lives on
GPU 0 (CUDA IPC peer mapping): the gather is the kernel's own pixel store over NVLink.  Strong
scaling: the frame is fixed, the ranks share it.  Timing: CUDA events, barrier + synchronize on both
sides, max over ranks.  Prints one JSON line (rank 0).

--check (rank 0, after the timed region): the gathered last frame must equal the same frame rendered
by GPU 0 alone, pixel for pixel; and a sample of its rays is searched by the CPU oracle (the
restatement, on the same flat tree): leaf ids and hit-point bits must match.
"""
import argparse
import ctypes as C
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

BANDS_PER_RANK = 8
STREAMS = 4
BYTES_INNER, BYTES_LEAF, BYTES_DENSE, BYTES_DOT, BYTES_PIXEL = 48, 32, 32, 16, 4


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--side", type=int, default=1024)
    ap.add_argument("--balls", type=int, default=4096)
    ap.add_argument("--w", type=int, default=7680)
    ap.add_argument("--h", type=int, default=4320)
    ap.add_argument("--frames", type=int, default=16)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=1)
    ap.add_argument("--check", action="store_true")
    ap.add_argument("--partition", default="interleave", choices=["interleave", "bands"],
                    help="interleave: rank r draws the 4-row tile rows r, r+N, ... in ONE launch per frame "
                         "(voxgpu_frame.interleave_*); bands: world*BANDS_PER_RANK row bands dealt round-robin")
    ap.add_argument("--simulate", type=int, default=0,
                    help="one GPU: time the share of each of N virtual ranks separately and report the balance")
    args = ap.parse_args()
    rank, local_rank, world = (int(os.environ.get(k, d)) for k, d in (("RANK", 0), ("LOCAL_RANK", 0), ("WORLD_SIZE", 1)))

    import voxvision_b200 as vv
    from voxvision_b200 import api, multi
    from tools import datasets as ds
    if not vv.gpu_available():
        raise SystemExit("c5_bench.py: no CUDA device")
    torch = dist = None
    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    vv.set_device(local_rank)
    g = api.libs.gpu
    W, H = args.w, args.h

    # ---- tree: built once, replicated ----------------------------------------------------------
    t0 = time.perf_counter()
    flat, n_vox, t_gen, t_build = None, 0, 0.0, 0.0
    if rank == 0:
        vox = ds.ball_union_voxels(args.side, args.balls)
        n_vox = len(vox)
        t_gen = time.perf_counter() - t0
        tree = vv.Tree.build(vox)
        del vox
        t_build = time.perf_counter() - t0 - t_gen
        if world > 1 or args.check:
            flat = tree.export()
    t1 = time.perf_counter()
    if world > 1:
        got = multi.broadcast_flat_tree(flat, 0, device=torch.device("cuda", local_rank))
        if rank != 0:
            tree = vv.Tree.from_arrays(got)
        del got
        torch.cuda.empty_cache()
    t_repl = time.perf_counter() - t1
    dev_tree = tree.device_handle()
    ti = tree.info()
    tree_bytes = int(g.voxgpu_tree_device_bytes(dev_tree))

    # ---- cameras: the C2 orbit scaled to the volume, `frames` equally spaced poses --------------
    scale = args.side / 256.0
    centre = (args.side / 2.0,) * 3
    cam = vv.Camera("simple-camera")
    gcams = []
    for k in range(args.frames):
        pos, look = ds.orbit_pose(k * 360.0 / args.frames, 360, centre, 328.0 * scale)
        ds.setup_camera(cam, pos, look, 1.2, W, H)
        gcams.append(cam.gpu_camera())
    ctx = vv.Context(16, 4)
    fmt, tex = ctx.pixfmt(), ctx.texture()
    d_tex = vv.DeviceBuffer.from_array(tex)

    # ---- the frame buffer on GPU 0, mapped by everyone -----------------------------------------
    frame_bytes = W * H * 4
    base, mode = None, "local"
    if world > 1:
        handle = torch.zeros(64, dtype=torch.uint8, device="cuda")
        if rank == 0:
            base = g.voxgpu_malloc(frame_bytes)
            hb = (C.c_ubyte * 64)()
            assert base and g.voxgpu_ipc_export(base, hb) == 0, g.voxgpu_last_error()
            handle.copy_(torch.tensor(list(hb), dtype=torch.uint8))
        dist.broadcast(handle, 0)
        if rank != 0:
            hb = (C.c_ubyte * 64)(*handle.cpu().tolist())
            base = g.voxgpu_ipc_open(hb)
        ok = torch.tensor([1 if base else 0], device="cuda")
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        if not int(ok.item()):
            raise SystemExit("c5_bench.py: peer mapping of GPU 0's frame buffer failed: " + g.voxgpu_last_error().decode())
        mode = "peer-store"
    else:
        own = vv.DeviceBuffer(frame_bytes)
        base = own.ptr
    bands = multi.interleaved_bands(H, world, BANDS_PER_RANK)[rank]
    streams = [C.c_void_p(g.voxgpu_stream_create()) for _ in range(STREAMS)]
    s0 = streams[0]

    def sync_all():
        if world > 1:
            torch.cuda.synchronize()
            dist.barrier()
            torch.cuda.synchronize()
        else:
            g.voxgpu_synchronize(None)

    def render_frame(k, vrank=rank, vworld=world):
        if args.partition == "interleave":
            fr = api.GpuFrame(W, H, 0, H, W, vv.QUALITY_BEST, fmt, None, vworld, vrank)
            rc = g.voxgpu_render(dev_tree, C.byref(gcams[k]), C.byref(fr), d_tex.ptr, base, None, None, streams[k % STREAMS])
            if rc:
                raise vv.VoxError(g.voxgpu_last_error().decode())
            return
        my = bands if vworld == world else multi.interleaved_bands(H, vworld, BANDS_PER_RANK)[vrank]
        for i, (y0, y1) in enumerate(my):
            fr = api.GpuFrame(W, H, y0, y1, W, vv.QUALITY_BEST, fmt, None)
            rc = g.voxgpu_render(dev_tree, C.byref(gcams[k]), C.byref(fr), d_tex.ptr, base + y0 * W * 4, None, None,
                                 streams[i % STREAMS])
            if rc:
                raise vv.VoxError(g.voxgpu_last_error().decode())

    def step():
        for st in streams[1:]:
            g.voxgpu_stream_join(st, s0)
        for k in range(args.frames):
            render_frame(k)
        for st in streams[1:]:
            g.voxgpu_stream_join(s0, st)

    timer = vv.Timer()
    if args.simulate > 1 and world == 1:
        # balance of an N-way split, measured on one GPU: each virtual rank's share of every frame, alone
        for part in ("interleave", "bands"):
            args.partition = part
            per_rank = []
            for vr in range(args.simulate):
                for k in range(args.frames):
                    render_frame(k, vr, args.simulate)
                g.voxgpu_synchronize(None)
                timer.start(s0)
                for st in streams[1:]:
                    g.voxgpu_stream_join(st, s0)
                for k in range(args.frames):
                    render_frame(k, vr, args.simulate)
                for st in streams[1:]:
                    g.voxgpu_stream_join(s0, st)
                timer.stop(s0)
                per_rank.append(timer.elapsed_ms())
            print(json.dumps({"simulate": args.simulate, "partition": part, "ms_per_rank": [round(x, 3) for x in per_rank],
                              "balance_mean_over_max": round(float(np.mean(per_rank) / np.max(per_rank)), 4)}), flush=True)
        return
    for _ in range(max(args.warmup, 1)):
        step()
    sync_all()
    timer.start(s0)
    for _ in range(args.steps):
        step()
    timer.stop(s0)
    sync_all()
    ms = timer.elapsed_ms()
    if world > 1:
        ms = multi.max_over_ranks(ms, device="cuda")

    # Frames of the timed region overlap on different streams and share the one frame buffer; for the
    # check every rank draws its share of the LAST pose once more, alone.
    render_frame(args.frames - 1)
    sync_all()

    # ---- checks and roofline bookkeeping (rank 0) ----------------------------------------------
    check = None
    bytes_per_ray = None
    if rank == 0:
        last = args.frames - 1
        got = np.empty((H, W), np.uint32)
        g.voxgpu_memcpy_d2h(got.ctypes.data, base, frame_bytes, None)
        g.voxgpu_synchronize(None)
        scratch = vv.DeviceBuffer(frame_bytes)
        fr = api.GpuFrame(W, H, 0, H, W, vv.QUALITY_BEST, fmt, None)
        d_leaf, d_pts = vv.DeviceBuffer(W * H * 4), vv.DeviceBuffer(W * H * 16)
        assert g.voxgpu_render(dev_tree, C.byref(gcams[last]), C.byref(fr), d_tex.ptr, scratch.ptr, d_leaf.ptr, d_pts.ptr, None) == 0
        alone = scratch.download(np.uint32, (H, W))
        check = {"gathered_frame_equals_single_gpu": bool(np.array_equal(got, alone)),
                 "lit_pixels": int((alone != 0).sum())}
        # visits per ray from the instrumented kernel (identical to the oracle's recursion,
        # tests/test_gpu_parity.py::test_traversal_counters_match_oracle) on every 4th row / column
        from oracle import orc
        P = orc.Oracle("port")
        pos, look = ds.orbit_pose(last * 360.0 / args.frames, 360, centre, 328.0 * scale)
        ocam = ds.setup_camera(P.camera(0), pos, look, 1.2, W, H)
        org, dirs = ocam.rays(W, H)
        sel = (np.arange(0, H, 4)[:, None] * W + np.arange(0, W, 4)[None, :]).reshape(-1)
        so, sd = np.ascontiguousarray(org[sel]), np.ascontiguousarray(dirs[sel])
        d_o, d_d = vv.DeviceBuffer.from_array(api.dots4(so)), vv.DeviceBuffer.from_array(api.dots4(sd))
        gs = api.GpuStats()
        assert g.voxgpu_rays_stats(dev_tree, 0, d_o.ptr, d_d.ptr, len(sel), None, None, None, C.byref(gs)) == 0
        st = gs.as_dict()
        bytes_per_ray = (st["inner"] * BYTES_INNER + st["leaf"] * BYTES_LEAF + st["dense"] * BYTES_DENSE +
                         st["dots_tested"] * BYTES_DOT) / len(sel) + BYTES_PIXEL
        if args.check:
            leaf = d_leaf.download(np.int32, (H * W,))
            pts = d_pts.download(np.float32, (H * W, 4))[:, :3]
            pick = sel[:: max(1, len(sel) // 65536)]
            otree = P.tree_from_flat(flat)
            oleaf, ohit, _ = otree.rays(org[pick], dirs[pick], threads=os.cpu_count() or 1)
            m = oleaf >= 0
            check["oracle_sample"] = {"rays": int(len(pick)), "hits": int(m.sum()),
                                      "leaf_ids_equal": bool(np.array_equal(leaf[pick], oleaf)),
                                      "hit_bits_equal": bool(np.array_equal(pts[pick][m].view(np.uint32), ohit[m].view(np.uint32)))}
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    rays = args.steps * args.frames * W * H
    peak = 6550.0
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            peak = float(json.load(f)["hbm_gbs"])
    except Exception:
        pass
    value = rays / (ms * 1e-3) / 1e6
    out = {
        "metric": "Mrays/s primary-ray render (C5 synthetic 1024^3, 7680x4320)", "value": round(value, 2), "unit": "Mrays/s",
        "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 1), "ms_per_step": round(ms / args.steps, 3),
        "ms_per_frame": round(ms / args.steps / args.frames, 4), "higher_is_better": True, "scaling": "strong", "dtype": "f32",
        "data": f"synthetic: union of {args.balls} balls with 1-in-8 hash holes in a {args.side}^3 grid, {n_vox} voxels",
        "config": {"workload": f"C5: {W}x{H}, {args.frames} poses of the scaled orbit, quality BEST, simple-camera",
                   "partition": (f"tile rows interleaved over {world} GPU(s) (rank r draws rows r, r+{world}, ...; one launch per frame)"
                                 if args.partition == "interleave" else
                                 f"{world * BANDS_PER_RANK} row bands dealt round-robin over {world} GPU(s)") +
                                f", tree replicated (built on rank 0, broadcast over NCCL), frame buffer on GPU 0 ({mode})",
                   "tree": {"nodes": int(ti.n_nodes), "inner": int(ti.n_inner), "leaf": int(ti.n_leaf), "dense": int(ti.n_dense),
                            "dots": int(ti.n_dots), "depth": int(ti.depth), "device_bytes": tree_bytes},
                   "host_seconds": {"generate": round(t_gen, 1), "vox_make_tree": round(t_build, 1), "replicate": round(t_repl, 1)},
                   "l2": "the device tree (GBs) exceeds the 126 MB L2; inputs are larger than L2"},
        "roofline": {"bound": "hbm", "bytes_per_ray": round(bytes_per_ray, 2),
                     "achieved": round(bytes_per_ray * value * 1e6 / 1e9 / world, 2), "peak": peak, "unit": "GB/s per GPU",
                     "frac": round(bytes_per_ray * value * 1e6 / 1e9 / world / peak, 5)},
        "check": check,
    }
    print(json.dumps(out), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
