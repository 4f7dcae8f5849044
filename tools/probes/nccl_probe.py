"""Probe: what does a rank-1 -> rank-0 transfer cost on this box (NCCL send/recv, gather, peer copy)?"""
import os, time, torch, torch.distributed as dist
rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); lr = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
x = torch.empty(64 << 20, dtype=torch.uint8, device="cuda")
outs = [torch.empty_like(x) for _ in range(world)] if rank == 0 else None
def timeit(f, n=10):
    f(); torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): f()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n
def sr():
    if rank == 1: dist.send(x, 0)
    elif rank == 0: dist.recv(x, 1)
t1 = timeit(sr); t2 = timeit(lambda: dist.gather(x, outs, dst=0)); t3 = timeit(lambda: dist.all_reduce(x.view(torch.float32)[: 16 << 20]))
if rank == 0:
    print(f"send/recv 64 MiB: {t1:.3f} ms = {64*1.048576/t1:.1f} GB/s; gather 64 MiB/rank: {t2:.3f} ms; all_reduce 64 MiB: {t3:.3f} ms")
    print("can access peer:", [torch.cuda.can_device_access_peer(0, i) for i in range(1, torch.cuda.device_count())])
dist.destroy_process_group()
