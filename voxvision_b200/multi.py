"""Multi-GPU plumbing for the render / query path: one process per GPU,
`torch.distributed` (NCCL on GPUs, gloo in the CPU tests) for the one exchange
step the path has -- bringing finished pixels (or query results) to rank 0.

The path shards without any data-path collective: rays, balls and pixels are
independent and the tree is read-only, so every rank holds a replica of the
tree (built and uploaded locally) and works on

  * its own frames of a camera path        (`frame_slice`, config C2), or
  * its own horizontal band of one frame   (`row_bands`, config C5; bands are
    multiples of 4 rows like the reference's 4x4 blocks, renderer.c:97-100), or
  * its own contiguous range of a batch    (`split_range`, configs C3/C4).

Nothing here touches CUDA directly; tensors live wherever the caller put them.
"""
import numpy as np
import torch
import torch.distributed as dist

FLAT_KEYS = (("bb_min", np.float32, 3), ("bb_max", np.float32, 3), ("flags", np.uint32, 1),
             ("dots_num", np.uint32, 1), ("center", np.float32, 3), ("first_child", np.uint32, 1),
             ("child_mask", np.uint32, 1), ("dots_off", np.uint64, 1), ("dots", np.float32, 3))


def broadcast_flat_tree(flat, src=0, device="cpu"):
    """Replicate a tree: `flat` (dict of arrays as Tree.export(), only needed on `src`) is broadcast to
    every rank -- over NCCL when `device` is a CUDA device, the arrays staged through it -- and returned
    as a dict of numpy arrays ready for Tree.from_arrays().  The tree is built ONCE (on src) instead of
    once per rank; for C5 that is ~3 GB of arrays against ~30 s of building."""
    rank = dist.get_rank()
    sizes = torch.zeros(3, dtype=torch.int64, device=device)
    if rank == src:
        sizes = torch.tensor([flat["flags"].shape[0], flat["dots"].shape[0], int(flat.get("depth", 0))],
                             dtype=torch.int64, device=device)
    dist.broadcast(sizes, src)
    n_nodes, n_dots, depth = (int(x) for x in sizes.tolist())
    out = {"depth": depth}
    for key, dtype, width in FLAT_KEYS:
        rows = n_dots if key == "dots" else n_nodes
        if rank == src:
            host = np.ascontiguousarray(flat[key], dtype).reshape(-1)
        else:
            host = np.empty(rows * width, dtype)
        # torch has no uint32/uint64 arithmetic needs here: ship the raw bytes
        t = torch.from_numpy(host.view(np.uint8))
        if str(device) != "cpu":
            d = t.to(device)
            dist.broadcast(d, src)
            if rank != src:
                t.copy_(d)
            del d
        else:
            dist.broadcast(t, src)
        out[key] = host.reshape(rows, width) if width > 1 else host
    return out


def split_range(n, rank, world):
    """[lo, hi) of rank's share of n independent items; shares differ by at most one."""
    base, extra = divmod(n, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def frame_slice(n_frames, rank, world):
    """Frames of a camera path rendered by `rank` (contiguous, so consecutive frames stay
    coherent for the frame-to-frame tile schedule)."""
    lo, hi = split_range(n_frames, rank, world)
    return range(lo, hi)


def row_bands(height, world):
    """[(y0, y1)] per rank: bands of whole 4-row units that tile [0, height)."""
    if height % 4:
        raise ValueError("frame height must be a multiple of 4 (renderer.c:97-100)")
    units = height // 4
    out = []
    for r in range(world):
        lo, hi = split_range(units, r, world)
        out.append((4 * lo, 4 * hi))
    return out


def interleaved_bands(height, world, bands_per_rank=4):
    """[[(y0, y1), ...] per rank]: the frame cut into world*bands_per_rank bands dealt round-robin,
    so that every rank gets a share of the expensive middle of the image (hit density is very
    uneven: 70-80% of a skull frame shows no voxel)."""
    cuts = row_bands(height, world * bands_per_rank)
    return [[b for i, b in enumerate(cuts) if i % world == r and b[1] > b[0]] for r in range(world)]


class FrameGather:
    """Gathers equally shaped per-rank tensors to `dst` (dist.gather), e.g. a batch of finished frames."""

    def __init__(self, like, dst=0):
        self.dst = dst
        self.rank = dist.get_rank()
        self.world = dist.get_world_size()
        self.slots = [torch.empty_like(like) for _ in range(self.world)] if self.rank == dst else None

    def __call__(self, local):
        dist.gather(local, self.slots, dst=self.dst)
        return self.slots


def gather_bands(local_rows, bands, width, dst=0):
    """Assemble one frame on `dst` from per-rank row bands.

    local_rows: (rows_of_this_rank, width) int32 tensor; bands: row_bands(height, world).
    Bands may differ by one 4-row unit, dist.gather wants equal shapes: pad to the tallest.
    Returns the (height, width) frame on dst, None elsewhere."""
    rank, world = dist.get_rank(), dist.get_world_size()
    tallest = max(y1 - y0 for y0, y1 in bands)
    send = local_rows
    if local_rows.shape[0] != tallest:
        send = torch.zeros((tallest, width), dtype=local_rows.dtype, device=local_rows.device)
        send[:local_rows.shape[0]] = local_rows
    slots = [torch.empty_like(send) for _ in range(world)] if rank == dst else None
    dist.gather(send.contiguous(), slots, dst=dst)
    if rank != dst:
        return None
    return torch.cat([slots[r][:bands[r][1] - bands[r][0]] for r in range(world)], dim=0)


def gather_ranges(local, n, dst=0):
    """Concatenate per-rank results of a split_range partition of n items on `dst`."""
    rank, world = dist.get_rank(), dist.get_world_size()
    sizes = [split_range(n, r, world) for r in range(world)]
    longest = max(hi - lo for lo, hi in sizes)
    send = local
    if local.shape[0] != longest:
        send = torch.zeros((longest,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
        send[:local.shape[0]] = local
    slots = [torch.empty_like(send) for _ in range(world)] if rank == dst else None
    dist.gather(send.contiguous(), slots, dst=dst)
    if rank != dst:
        return None
    return torch.cat([slots[r][:sizes[r][1] - sizes[r][0]] for r in range(world)], dim=0)


def max_over_ranks(value, device="cpu"):
    """Max of a python float over all ranks (device timings are reported as the slowest rank's)."""
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())
