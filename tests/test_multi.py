"""N > 1 host logic with world_size 2 on gloo (CPU): partitions tile the work exactly, the
gathers reassemble on rank 0 what one process would have produced."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from voxvision_b200 import multi


def test_partitions_tile_the_work():
    for n in (0, 1, 7, 360, 1 << 24):
        for world in (1, 2, 3, 8):
            parts = [multi.split_range(n, r, world) for r in range(world)]
            assert parts[0][0] == 0 and parts[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(parts, parts[1:]))
            sizes = [hi - lo for lo, hi in parts]
            assert max(sizes) - min(sizes) <= 1
    for h in (4, 600, 1080, 4320):
        for world in (1, 2, 4, 8):
            bands = multi.row_bands(h, world)
            assert bands[0][0] == 0 and bands[-1][1] == h
            assert all(a[1] == b[0] for a, b in zip(bands, bands[1:]))
            assert all(y0 % 4 == 0 and y1 % 4 == 0 for y0, y1 in bands)
            inter = multi.interleaved_bands(h, world)
            rows = sorted(b for r in inter for b in r)
            assert rows[0][0] == 0 and rows[-1][1] == h and all(a[1] == b[0] for a, b in zip(rows, rows[1:]))
    with pytest.raises(ValueError):
        multi.row_bands(1081, 2)
    assert list(multi.frame_slice(360, 1, 8)) == list(range(45, 90))


def fake_frame(h, w, y0, y1, frame):
    """Stands in for the render kernel on a CPU box: a deterministic function of (frame, x, y)."""
    y, x = np.mgrid[y0:y1, 0:w]
    return torch.from_numpy(((frame * 2654435761 + y * 40503 + x * 9973) & 0x7fffffff).astype(np.int32))


def worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        h, w = 36, 48                                    # 9 four-row units: uneven over 2 ranks
        bands = multi.row_bands(h, world)
        y0, y1 = bands[rank]
        frame = multi.gather_bands(fake_frame(h, w, y0, y1, 7), bands, w)
        if rank == 0:
            assert torch.equal(frame, fake_frame(h, w, 0, h, 7))
        else:
            assert frame is None

        # frames of a path, gathered in batches like bench.py does
        mine = multi.frame_slice(10, rank, world)
        batch = torch.stack([fake_frame(8, 16, 0, 8, f) for f in mine])
        got = multi.FrameGather(batch)(batch)
        if rank == 0:
            for r in range(world):
                want = torch.stack([fake_frame(8, 16, 0, 8, f) for f in multi.frame_slice(10, r, world)])
                assert torch.equal(got[r], want)

        # a ray / ball batch split into contiguous ranges
        n = 1001
        lo, hi = multi.split_range(n, rank, world)
        res = multi.gather_ranges(torch.arange(lo, hi, dtype=torch.int64) * 3, n)
        if rank == 0:
            assert torch.equal(res, torch.arange(n, dtype=torch.int64) * 3)

        # tree replication: built on rank 0, broadcast as flat arrays, imported everywhere
        import voxvision_b200 as vv
        from tools import datasets as ds
        vox = ds.random_voxels(side=24, p=0.2)[0]
        flat = vv.Tree.build(vox).export() if rank == 0 else None
        got = multi.broadcast_flat_tree(flat, src=0)
        mine = vv.Tree.build(vox).export()
        for k, _, _ in multi.FLAT_KEYS:
            assert np.array_equal(np.asarray(got[k]).view(np.uint8), np.asarray(mine[k]).view(np.uint8)), k
        assert vv.Tree.from_arrays(got).voxels() == len(vox)

        slowest = multi.max_over_ranks(1.0 + rank)
        assert slowest == float(world)
        open(os.path.join(out_dir, f"ok{rank}"), "w").close()
    finally:
        dist.destroy_process_group()


def test_world_size_2_gloo(tmp_path):
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mp.spawn(worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    assert os.path.exists(tmp_path / "ok0") and os.path.exists(tmp_path / "ok1")


def test_band_of_a_tile_row_is_the_inverse_of_the_band_cuts():
    """The presenter cuts a frame of `units` 4-row tile rows into n bands at (units*i)//n (present.inl); k_list
    (render_launch.inl) finds the band of tile row r as ((r+1)*n - 1)//units.  The two must agree for every
    geometry, or a band's list would receive tiles of its neighbour."""
    for units in list(range(1, 40)) + [135, 270, 1080, 1081]:
        for n in range(1, min(units, 16) + 1):
            cuts = [units * i // n for i in range(n + 1)]
            assert cuts[0] == 0 and cuts[-1] == units
            for r in range(units):
                b = ((r + 1) * n - 1) // units
                assert cuts[b] <= r < cuts[b + 1], (units, n, r, b)


def test_interleaved_shares_partition_the_tile_rows():
    """voxgpu_frame.interleave_*: rank k of n draws the tile rows r with r % n == k; the number of rows the
    kernel computes for itself, (rows - k + n - 1)//n, is the size of that share, and the shares tile the frame."""
    for rows in (1, 2, 7, 135, 270, 1080):
        for n in (1, 2, 3, 4, 8, 16):
            seen = []
            for k in range(n):
                own = (rows - k + n - 1) // n if k < rows else 0
                share = list(range(k, rows, n))
                assert own == len(share)
                seen += share
            assert sorted(seen) == list(range(rows))
