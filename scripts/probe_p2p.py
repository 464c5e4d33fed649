"""Development probe (round 2): how can N ranks exchange the slabs of their partial grids on this box?
 (a) torch symmetric memory (rendezvous + peer views), (b) legacy CUDA IPC handles + cudaMemcpyAsync pulls on the copy
 engines, (c) NCCL collectives on the same payloads.  Times are CUDA events on the issuing stream, max over ranks.
 usage: torchrun --nproc-per-node N scripts/probe_p2p.py"""
import ctypes, os, sys, time
import torch
import torch.distributed as dist

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
dev = torch.device("cuda", local)
S = 129 * 16384 * 4            # elements of the interleaved cube accumulator (float64): 67.6 MB
chunk = (S + world - 1) // world

def say(*a):
    if rank == 0:
        print(*a, flush=True)

def timed(fn, reps=10, warm=3):
    for _ in range(warm): fn()
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): fn()
    b.record(); torch.cuda.synchronize()
    t = torch.tensor([a.elapsed_time(b) / reps], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())

# ---- (c) NCCL collectives
x = torch.ones(chunk * world, dtype=torch.float64, device=dev)
y = torch.empty(chunk, dtype=torch.float64, device=dev)
z = torch.empty(chunk * world, dtype=torch.float64, device=dev)
tiny = torch.ones(1, device=dev)
say("NCCL tiny all_reduce            %.4f ms" % timed(lambda: dist.all_reduce(tiny), 50))
say("NCCL reduce           %5.1f MB  %.4f ms" % (x.numel() * 8 / 1e6, timed(lambda: dist.reduce(x, dst=0))))
say("NCCL reduce_scatter   %5.1f MB  %.4f ms" % (x.numel() * 8 / 1e6, timed(lambda: dist.reduce_scatter_tensor(y, x))))
say("NCCL all_to_all       %5.1f MB  %.4f ms" % (x.numel() * 8 / 1e6, timed(lambda: dist.all_to_all_single(z, x))))
h = x[: x.numel() // 2]
yh = torch.empty(h.numel() // world, dtype=torch.float64, device=dev)
say("NCCL reduce           %5.1f MB  %.4f ms" % (h.numel() * 8 / 1e6, timed(lambda: dist.reduce(h, dst=0))))
say("NCCL reduce_scatter   %5.1f MB  %.4f ms" % (h.numel() * 8 / 1e6, timed(lambda: dist.reduce_scatter_tensor(yh, h[: yh.numel() * world]))))
say("NCCL all_gather       %5.1f MB  %.4f ms" % (z.numel() * 8 / 1e6, timed(lambda: dist.all_gather_into_tensor(z, y))))

# ---- (b) legacy CUDA IPC + copy-engine pulls
try:
    rt = ctypes.CDLL("libcudart.so.12")
    class Handle(ctypes.Structure):
        _fields_ = [("reserved", ctypes.c_char * 64)]
    rt.cudaIpcGetMemHandle.argtypes = [ctypes.POINTER(Handle), ctypes.c_void_p]
    rt.cudaIpcOpenMemHandle.argtypes = [ctypes.POINTER(ctypes.c_void_p), Handle, ctypes.c_uint]
    rt.cudaMalloc.argtypes = [ctypes.POINTER(ctypes.c_void_p), ctypes.c_size_t]
    rt.cudaMemcpyAsync.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int, ctypes.c_void_p]
    buf = ctypes.c_void_p()
    assert rt.cudaMalloc(ctypes.byref(buf), chunk * world * 8) == 0
    hdl = Handle()
    rc = rt.cudaIpcGetMemHandle(ctypes.byref(hdl), buf)
    assert rc == 0, "cudaIpcGetMemHandle rc=%d" % rc
    mine = torch.tensor(list(bytes(hdl.reserved) if isinstance(hdl.reserved, bytes) else ctypes.string_at(ctypes.byref(hdl), 64)), dtype=torch.uint8, device=dev)
    mine = torch.tensor(list(ctypes.string_at(ctypes.byref(hdl), 64)), dtype=torch.uint8, device=dev)
    allh = torch.empty(world * 64, dtype=torch.uint8, device=dev)
    dist.all_gather_into_tensor(allh, mine)
    allh = allh.cpu().numpy().tobytes()
    peers = []
    for r in range(world):
        if r == rank:
            peers.append(buf.value); continue
        hr = Handle()
        ctypes.memmove(ctypes.byref(hr), allh[r * 64:(r + 1) * 64], 64)
        p = ctypes.c_void_p()
        rc = rt.cudaIpcOpenMemHandle(ctypes.byref(p), hr, 1)
        assert rc == 0, "cudaIpcOpenMemHandle rc=%d" % rc
        peers.append(p.value)
    say("CUDA IPC: handles opened on every rank")
    stage = torch.empty(chunk * world, dtype=torch.float64, device=dev)
    streams = [torch.cuda.Stream() for _ in range(world)]
    def pull(nstreams):
        cur = torch.cuda.current_stream()
        ev = torch.cuda.Event(); ev.record(cur)
        for k in range(1, world):
            r = (rank + k) % world
            st = streams[k % nstreams]
            st.wait_event(ev)
            rc = rt.cudaMemcpyAsync(stage.data_ptr() + r * chunk * 8, peers[r] + rank * chunk * 8, chunk * 8, 4, ctypes.c_void_p(st.cuda_stream))
            assert rc == 0
        for st in streams[:nstreams]:
            e2 = torch.cuda.Event(); e2.record(st); cur.wait_event(e2)
    for ns in (1, 2, 4, 7):
        ns = min(ns, max(1, world - 1))
        say("IPC pull of %d x %.1f MB on %d stream(s)   %.4f ms" % (world - 1, chunk * 8 / 1e6, ns, timed(lambda: pull(ns))))
    # the same pulls while a long compute kernel owns the SMs
    a = torch.randn(8192, 8192, device=dev, dtype=torch.bfloat16)
    def busy_pull():
        side = streams[-1]
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(4): a @ a
        pull(min(4, max(1, world - 1)))
        torch.cuda.current_stream().wait_stream(side)
    def busy():
        for _ in range(4): a @ a
    tb, tbp = timed(busy), timed(busy_pull)
    say("4 GEMMs alone %.4f ms; with the pulls next to them %.4f ms" % (tb, tbp))
except Exception as e:
    say("CUDA IPC path failed:", repr(e))

# ---- (a) torch symmetric memory
try:
    import torch.distributed._symmetric_memory as symm
    t = symm.empty(chunk * world, dtype=torch.float64, device=dev)
    hd = symm.rendezvous(t, dist.group.WORLD.group_name)
    t.fill_(rank + 1.0)
    hd.barrier()
    peer = hd.get_buffer((rank + 1) % world, (chunk * world,), torch.float64)
    got = float(peer[0].item())
    say("symmetric memory: rendezvous ok, peer value %.1f, multicast_ptr %s" % (got, getattr(hd, "multicast_ptr", None)))
    say("symm barrier   %.4f ms" % timed(lambda: hd.barrier(), 50))
except Exception as e:
    say("symmetric memory failed:", repr(e)[:300])
dist.barrier()
dist.destroy_process_group()
