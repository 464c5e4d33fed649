"""Development: which collective sums a 34 MB float64 / int64 grid fastest across the ranks?"""
import os, torch, torch.distributed as dist
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
n = 128 * 128 * 256 + 128 * 128 * 4
def timeit(fn, reps=20):
    for _ in range(5): fn()
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): fn()
    b.record(); torch.cuda.synchronize()
    t = torch.tensor([a.elapsed_time(b) / reps], device="cuda"); dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return t.item()
for dt in (torch.float64, torch.int64, torch.float32):
    x = torch.ones(n, dtype=dt, device="cuda")
    pad = (-n) % world
    xs = torch.ones(n + pad, dtype=dt, device="cuda")
    out = torch.empty((n + pad) // world, dtype=dt, device="cuda")
    r = {"reduce": timeit(lambda: dist.reduce(x, 0)), "all_reduce": timeit(lambda: dist.all_reduce(x)),
         "reduce_scatter": timeit(lambda: dist.reduce_scatter_tensor(out, xs))}
    if rank == 0:
        print(dt, " ".join("%s=%.3f ms" % kv for kv in r.items()), flush=True)
dist.destroy_process_group()
