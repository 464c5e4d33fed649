#!/usr/bin/env python
"""bench.py -- particles/s binned to moment maps + LOSVD cube (BASELINE.json metric).

Workload (BASELINE.json configs[1], SURVEY.md 8d "C2"): a seeded synthetic disk+bulge snapshot of
10^8 float32 particles PER GPU (weak scaling), rotate(PA=30, incl=60) -> 128x128 mass / V / sigma
maps + 128x128x256 LOSVD cube from ONE read of the particles -> Gaussian smoothing of the cube
(convolve_spatial to FWHM 1 kpc).  At N > 1 every rank bins its own shard and the partial grids are
summed with an NCCL reduce to the step's owner rank (step k -> rank k mod N), which finalises and smooths.

  python bench.py [--gpus N] [--steps K] [--warmup W]            our arm (default N=1, K=30, W=5)
  python bench.py --impl reference [--gpus N] [--steps K] ...    the reference's CPU path (numpy)

A "step" is one full pass of the path over the resident particles.  One JSON line on stdout.
"""
import argparse
import contextlib
import io
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

N_PER_GPU = 100_000_000
NXY, NV = 128, 256
PA, INCL = 30.0, 60.0
SMOOTH_FWHM = (1.0, 1.0)
BYTES_PER_PARTICLE = 28          # x y z vx vy vz w, float32
SEED = 1234 + 1                  # base seed + config index
CPU_SAMPLE = int(os.environ.get("PNM_CPU_SAMPLE", 4_000_000))   # particles of the bounded CPU sample
METRIC = "particles/s binned to maps+LOSVD cube"


def workload_name(n_gpus, n_per_gpu):
    return ("%.0e-particle disk+bulge per GPU (x%d), rotate(PA=30,incl=60), 128x128 M/V/sigma maps + "
            "128x128x256 LOSVD cube + Gaussian smoothing (FWHM 1 kpc)" % (n_per_gpu, n_gpus))


# ------------------------------------------------------------------------------------------ CPU arm
_SAMPLE = None      # (7, n) float32 sample, inherited by forked workers (no pickling of the inputs)


def _cpu_bin_slice(bounds):
    """Worker: the reference's numpy call sequence on one slice (rotation.py:115-117,
    grid.py:126-131, grid.py:252-259 through oracle_np.project_path)."""
    from oracle import oracle_np as on
    import workloads
    a, b = bounds
    pos = np.ascontiguousarray(_SAMPLE[0:3, a:b].T)
    vel = np.ascontiguousarray(_SAMPLE[3:6, a:b].T)
    mass = np.ascontiguousarray(_SAMPLE[6, a:b])
    try:
        from threadpoolctl import threadpool_limits
        ctx = threadpool_limits(1) if os.environ.get("PNM_CPU_WORKER") else contextlib.nullcontext()
    except Exception:
        ctx = contextlib.nullcontext()
    with ctx:
        s0, V, S, cube = on.project_path(pos, vel, mass, PA, INCL, workloads.LIM_XY, NXY, workloads.LIM_V, NV)
    return s0, cube


def _worker_init():
    os.environ["PNM_CPU_WORKER"] = "1"


def cpu_reference_step(pool, nproc):
    """One bounded-sample step of the reference CPU path, parallelised over `nproc` host processes
    by particle slices (the reference itself is single threaded apart from np.dot).  Returns
    (seconds binning the sample, seconds smoothing the cube)."""
    from oracle import oracle_c as oc
    from oracle import oracle_np as on
    import workloads
    n = _SAMPLE.shape[1]
    edges = np.linspace(0, n, nproc + 1).astype(np.int64)
    jobs = list(zip(edges[:-1].tolist(), edges[1:].tolist()))
    t0 = time.perf_counter()
    parts = pool.map(_cpu_bin_slice, jobs) if pool is not None else [_cpu_bin_slice(j) for j in jobs]
    cube = parts[0][1]
    for p in parts[1:]:
        cube = cube + p[1]
    t_bin = time.perf_counter() - t0
    step = (workloads.LIM_XY[1] - workloads.LIM_XY[0]) / (NXY - 1)
    k2 = on.gaussian_kernel2d(SMOOTH_FWHM[0] * on.FWHM_TO_SIGMA / step, SMOOTH_FWHM[1] * on.FWHM_TO_SIGMA / step)
    t0 = time.perf_counter()
    oc.smooth_cube(cube, k2)        # losvd.py:127-135 restated in C (astropy's convolve is a C loop too)
    t_smooth = time.perf_counter() - t0
    return t_bin, t_smooth


def cpu_baseline(steps, warmup, n_workload, sample_n=CPU_SAMPLE):
    """Times the oracle's restatement of the reference path on the host cores.  The sample is the
    first `sample_n` particles of the same synthetic workload; the figure is extrapolated to the
    whole workload as n / (t_bin * n / sample + t_smooth) because binning is linear in n
     and the smoothing cost does not depend on n."""
    import multiprocessing as mp
    import workloads
    global _SAMPLE
    nproc = os.cpu_count() or 1
    _SAMPLE = workloads.make_snapshot(sample_n, seed=SEED, device="cpu").numpy()
    # fork: workers inherit the sample; this runs before any CUDA initialisation in the process
    pool = mp.get_context("fork").Pool(nproc, initializer=_worker_init) if nproc > 1 else None
    try:
        times = []
        for i in range(warmup + steps):
            tb, ts = cpu_reference_step(pool, nproc)
            if i >= warmup:
                times.append((tb, ts))
    finally:
        if pool is not None:
            pool.close()
            pool.join()
    tb = float(np.mean([t[0] for t in times]))
    ts = float(np.mean([t[1] for t in times]))
    t_full = tb * n_workload / sample_n + ts
    return {
        "value": n_workload / t_full, "unit": "particles/s", "cores": nproc, "kind": "port",
        "sample": ("first %d particles of the workload split over %d processes: numpy np.dot + histogram2d x3 + "
                   "histogramdd (%.2f s) + C direct 15x15 convolution of the 128x128x256 cube (%.2f s); "
                   "extrapolated linearly in n to %d particles" % (sample_n, nproc, tb, ts, n_workload)),
        "t_bin_sample_s": tb, "t_smooth_s": ts,
    }, (tb * n_workload / sample_n + ts)


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    n_workload = args.particles * args.gpus
    steps, warmup = max(1, args.steps), max(0, args.warmup)
    # bounded: each step is ~2-4 s of wall clock; cap so that the run ends within a few minutes
    steps_run, warmup_run = min(steps, 20), min(warmup, 2)
    base, t_step = cpu_baseline(steps_run, warmup_run, n_workload)
    sample_ms = (base["t_bin_sample_s"] + base["t_smooth_s"]) * 1e3
    line = {
        "impl": "reference", "metric": METRIC, "value": base["value"], "unit": "particles/s",
        "n_gpus": args.gpus, "steps": steps_run, "warmup": warmup_run,
        # what one timed step really took: the bounded sample (binning of CPU_SAMPLE particles + smoothing of the cube);
        # `value` is the whole-workload throughput extrapolated from it (binning is linear in n, smoothing constant)
        "ms_per_step": sample_ms, "ms_per_step_whole_workload_extrapolated": t_step * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_name(args.gpus, args.particles), "particles_per_gpu": args.particles, "nXY": NXY, "nV": NV,
                   "accumulate": "f64",
                   "accumulators": "float64 (numpy histogram2d / histogramdd)",
                   "smoothing_parity": "unpinned (astropy absent: oracle restates convolve() from its documentation)",
                   "parallelism": "%d host processes over particle slices of a %d-particle sample (the reference itself is single threaded)"
                                  % (base["cores"], CPU_SAMPLE),
                   "l2": "n/a (CPU arm)", "streams": "n/a (CPU arm)"},
        "note": "CPU port of the reference path (pure-Python reference: its numpy call sequence restated in oracle/); "
                "requested steps/warmup %d/%d capped to %d/%d so that the run ends within minutes" % (steps, warmup, steps_run, warmup_run),
        "cpu_baseline": base,
        "e2e": {"value": base["value"], "unit": "particles/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))
    return 0


# ------------------------------------------------------------------------------------------ clocks
class ClockSampler(object):
    """Samples the SM clock and the throttle reasons of one GPU while the timed region runs -- from a CHILD PROCESS
    (`nvidia-smi --query-gpu ... -lms 20`, plus one NVML read by the launching thread while the queue drains), not a thread: a polling thread shares the GIL with the launching loop, and in
    a sharded run a launch loop that loses a few milliseconds stalls every rank at the next device-side barrier."""
    REASONS = {0x4: "sw_power_cap", 0x8: "hw_slowdown", 0x20: "sw_thermal_slowdown", 0x40: "hw_thermal_slowdown",
               0x80: "hw_power_brake_slowdown", 0x2: "applications_clocks_setting", 0x10: "sync_boost"}

    def __init__(self, index, period_ms=20):
        import shutil
        import subprocess
        self.proc, self.ok, self.first, self.extra, self.nvml = None, False, "", [], None
        try:                                                 # NVML handle for sample_now() (no polling thread)
            import pynvml
            pynvml.nvmlInit()
            self.nvml = (pynvml, pynvml.nvmlDeviceGetHandleByIndex(index))
        except Exception:
            self.nvml = None
        exe = shutil.which("nvidia-smi")
        if exe is None:
            return
        try:
            self.proc = subprocess.Popen(
                [exe, "-i", str(index), "--query-gpu=clocks.sm,clocks.max.sm,clocks_event_reasons.active",
                 "--format=csv,noheader,nounits", "-lms", str(period_ms)],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.first = self.proc.stdout.readline()         # blocks until the child is polling (driver start-up)
            self.ok = bool(self.first.strip())
        except Exception:
            self.proc = None

    def sample_now(self):
        """One NVML sample taken by the launching thread itself (called right after the last launch of the timed
        region, while the GPU is still working through the queue)."""
        if self.nvml is not None:
            try:
                nv, h = self.nvml
                get = getattr(nv, "nvmlDeviceGetCurrentClocksEventReasons", None) or nv.nvmlDeviceGetCurrentClocksThrottleReasons
                self.extra.append((float(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)), int(get(h))))
            except Exception:
                pass

    def stop(self):
        out = ""
        if self.proc is not None:
            self.proc.terminate()
            try:
                out, _ = self.proc.communicate(timeout=5)
            except Exception:
                self.proc.kill()
        samples, mask, max_mhz = [c for c, _ in self.extra], 0, None
        for _, m in self.extra:
            mask |= m
        lines = out.splitlines() if self.ok else []
        for line in (lines if lines else ([self.first] if self.ok else [])):      # (the first sample was taken before the timed region)
            parts = [p.strip() for p in line.split(",")]
            try:
                samples.append(float(parts[0]))
                max_mhz = float(parts[1])
                mask |= int(parts[2], 16)
            except Exception:
                continue
        if max_mhz is None and self.nvml is not None:
            try:
                max_mhz = float(self.nvml[0].nvmlDeviceGetMaxClockInfo(self.nvml[1], self.nvml[0].NVML_CLOCK_SM))
            except Exception:
                pass
        reasons = [name for bit, name in self.REASONS.items() if mask & bit]
        return {"sm_mhz": float(np.median(samples)) if samples else None, "sm_max_mhz": max_mhz,
                "reasons": reasons, "samples": len(samples), "sampler": "nvidia-smi -lms 20 (child process) + one NVML read after the last launch"}


# ------------------------------------------------------------------------------------------ our arm
BIG_N = 1_000_000_000


def _sha(t):
    import hashlib
    return hashlib.sha256(t.detach().contiguous().cpu().numpy().tobytes()).hexdigest()[:16]


def parity_check(pnm, engine, workloads, snap, soa, reducer, rank, world, n_local, sink):
    """One step with int64 fixed-point accumulators through the bench's own path (k_project_cube + the reduce over
    ranks); rank 0 then re-bins every rank's shard sequentially into ONE accumulator with the same scales and
    compares the finalised maps and cube bit for bit."""
    import torch
    from pynmap_b200 import _lib
    kw = dict(weight_name="mass", limXY=workloads.LIM_XY, nXY=NXY, limV=workloads.LIM_V, nV=NV, accumulate="fixed")
    grid = engine.get_grid(workloads.LIM_XY, NXY, NXY, workloads.LIM_V, NV)
    if reducer is not None:
        reducer.dst = 0
    with contextlib.redirect_stdout(sink):
        vm, lv = snap.project(reducer=reducer, **kw)
    if getattr(reducer, "scatter_slabs", False):         # every rank holds its own velocity bins: assemble them on rank 0
        whole = reducer.gather_losvd(lv.losvd, grid, dst=0)
        if rank == 0:
            lv.losvd = whole
    torch.cuda.synchronize()
    scales = snap.last_fixed_scales
    if rank != 0:
        return None
    acc = engine.AccMode("fixed")
    sums = 31 | _lib.CUBE_MOMENTS
    _, cube = engine.new_accumulators(grid, sums, acc)
    mat = pnm.rotation.rotation_matrix(np.float32(PA), np.float32(INCL))
    half = (n_local // 2 + 1029) // 8 * 8
    pieces = []
    for r in range(world):
        t = soa if r == 0 else workloads.make_snapshot(n_local, bulge_fraction=0.2, seed=SEED + 1000 * r, device="cuda")
        soa_r = [t[k] for k in range(7)]
        cuts = [(0, n_local)] if world > 1 else [(0, half), (half, n_local)]
        for lo, hi in cuts:
            rows = [a[lo:hi] for a in soa_r]
            plan = engine.CubePlan(grid).build(rows[:6], rows[6], mat)
            engine.project_cube(grid, plan, rows[:6], rows[6], mat, acc, scales, cube)
            pieces.append(hi - lo)
        torch.cuda.synchronize()
        del soa_r, t
    M, V, S = engine.finalize_vmaps(grid, None, cube, sums, acc, scales)
    c = engine.finalize_cube(grid, cube, acc, scales, sums)
    same = lambda a, b: bool(torch.equal(torch.nan_to_num(a, nan=-1.0), torch.nan_to_num(b, nan=-1.0)))   # noqa: E731
    equal = same(vm.M, M) and same(vm.V, V) and same(vm.S, S) and same(lv.losvd, c)
    return {"checked": True, "n_ranks": world, "accumulate": "fixed", "kernel": "k_project_cube",
            "sha256_maps": _sha(torch.stack([vm.M, vm.V, vm.S])), "sha256_cube": _sha(lv.losvd),
            "sha256_maps_sequential": _sha(torch.stack([M, V, S])), "sha256_cube_sequential": _sha(c),
            "equals_1rank": equal, "particles_binned": float(vm.M.sum().item()),
            "how": ("rank 0 re-binned the %d shards one after the other into one accumulator (same scales)" % world) if world > 1
                   else "the snapshot in one launch vs two ragged half shards (%d + %d particles) into one accumulator" % tuple(pieces)}


def c3_point(pnm, workloads, reducer, rank, world, sink, n_total=1_000_000_000, nxy=512):
    """BASELINE.json configs[2]: n_total particles sharded over the ranks, nxy x nxy M / V / sigma maps, partial grids
    combined by an NCCL all-reduce.  Single passes (no pipelining), timed with CUDA events, max over ranks."""
    import torch
    import torch.distributed as dist
    n_local = n_total // world
    soa = workloads.make_snapshot(n_local, bulge_fraction=0.2, seed=SEED + 2 + 1000 * rank, device="cuda")
    snap = pnm.snapshot.from_arrays(soa[:3], soa[3:6], mass=soa[6], output="torch")
    del soa
    snap.rotate(PA=PA, inclination=INCL)

    def one():
        with contextlib.redirect_stdout(sink):
            vm = snap.velocity_moments(weight_name="mass", limXY=workloads.LIM_XY, nXY=nxy, reducer=reducer)
        sink.seek(0); sink.truncate()
        return vm

    for _ in range(3):
        one()
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    reps = 10
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        vm = one()
    b.record(); torch.cuda.synchronize()
    ms = torch.tensor([a.elapsed_time(b) / reps], dtype=torch.float64, device="cuda")
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms = float(ms.item())
    binned = float(vm.M.sum().item())
    del snap
    torch.cuda.empty_cache()
    return {"workload": "%.3g particles over %d GPUs (%.3g per GPU), %dx%d M/V/sigma maps, all-reduce of the partial grids; "
                        "one non-pipelined pass" % (n_total, world, n_local, nxy, nxy),
            "ms_per_pass": ms, "value": n_local * world / (ms * 1e-3), "unit": "particles/s", "particles_binned": binned}


def big_point(pnm, engine, workloads, n, peak, sink, accumulate):
    """10^9 particles resident on ONE GPU: the same pipelined step, 5 timed steps after 3 warm-up steps."""
    import torch
    soa = workloads.make_snapshot(n, bulge_fraction=0.2, seed=SEED + 7, device="cuda")
    snap = pnm.snapshot.from_arrays(soa[:3], soa[3:6], mass=soa[6], output="torch")
    del soa
    snap.rotate(PA=PA, inclination=INCL)
    post = torch.cuda.Stream(priority=-1)

    def step():
        with contextlib.redirect_stdout(sink):
            _, lv = snap.project(weight_name="mass", limXY=workloads.LIM_XY, nXY=NXY, limV=workloads.LIM_V, nV=NV,
                                 accumulate=accumulate, post_stream=post)
            with torch.cuda.stream(post):
                lv.convolve_spatial(*SMOOTH_FWHM)
        sink.seek(0); sink.truncate()

    for _ in range(3):
        step()
    torch.cuda.synchronize()
    engine.KERNEL_TIMER = timer = engine.KernelTimer()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(5):
        step()
    torch.cuda.current_stream().wait_stream(post)
    b.record(); torch.cuda.synchronize()
    engine.KERNEL_TIMER = None
    ms, kms = a.elapsed_time(b) / 5, timer.total_ms() / max(1, len(timer.pairs))
    gbs = BYTES_PER_PARTICLE * n / (kms * 1e-3) / 1e9
    return {"particles": n, "steps": 5, "ms_per_step": ms, "value": n / (ms * 1e-3), "unit": "particles/s",
            "kernel": timer.kernel, "kernel_ms": kms, "achieved_gbs": gbs, "roofline_frac": gbs / peak}


def run_b200(args):
    import torch
    import torch.distributed as dist

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    # stdout carries ONE JSON line: native libraries (NCCL prints its version banner on the C stdout) are sent to stderr
    sys.stdout.flush()
    json_out = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    if world != args.gpus and world > 1:
        raise SystemExit("--gpus %d but WORLD_SIZE=%d" % (args.gpus, world))
    n_gpus = world
    n_local = args.particles

    base = None
    if rank == 0 and n_gpus == 1 and not args.no_cpu_baseline:
        base, _ = cpu_baseline(1, 0, n_local)          # before CUDA is initialised in this process

    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        opts = dist.ProcessGroupNCCL.Options(is_high_priority_stream=True)   # the reduce must not queue behind the scatter
        if args.nccl_max_ctas > 0:               # fewer reduce CTAs = fewer SMs taken from the scatter kernel
            opts.config.max_ctas = args.nccl_max_ctas
            opts.config.min_ctas = min(opts.config.min_ctas, args.nccl_max_ctas) if opts.config.min_ctas > 0 else 1
        dist.init_process_group("nccl", device_id=torch.device("cuda", local), pg_options=opts)

    import pynmap_b200 as pnm
    import workloads
    from pynmap_b200 import engine, hostpath
    from pynmap_b200.sharded import GridReducer, PeerSlabReducer, SlabReducer

    pnm._lib.require_device()
    soa = workloads.make_snapshot(n_local, bulge_fraction=0.2, seed=SEED + 1000 * rank, device="cuda")
    snap = pnm.snapshot.from_arrays(soa[:3], soa[3:6], mass=soa[6], output="torch")
    snap.rotate(PA=PA, inclination=INCL)
    reducer = e2e_reducer = None
    if world > 1:
        # the scatter kernel is persistent (one CTA per SM): what runs next to it needs SMs of its own
        reserve = args.reserve_sms if args.reserve_sms >= 0 else 0
        reducer = (GridReducer(reserve_sms=reserve) if args.reduce == "owner" else
                   SlabReducer(reserve_sms=reserve) if args.reduce == "scatter" else PeerSlabReducer())
        e2e_reducer = GridReducer()
    sink = io.StringIO()

    # post-processing stream: replica fold, NCCL reduce, finalisation and smoothing of step k run
    # next to the scatter kernel of step k+1 (all K steps are complete before the timer stops)
    post = None if args.no_pipeline else torch.cuda.Stream(priority=-1)

    counter = [0]
    deferred = world > 1 and args.reduce == "peer"
    if deferred:
        post = None                              # software pipelining on ONE stream instead (see step())
    pending = [None]

    def finish(p):
        with contextlib.redirect_stdout(sink):
            vmap, losvd = p.result()
            losvd = losvd.convolve_spatial(*SMOOTH_FWHM)
        sink.seek(0)
        sink.truncate()
        return vmap, losvd

    def drain():
        if pending[0] is not None:
            finish(pending[0])
            pending[0] = None

    def step():
        if deferred:
            # scatter of step k, then the finalisation of step k - 1 whose chunks were pulled from the peers by the copy
            # engines while the scatter kernels ran: all on one stream, nothing competes with the persistent kernel for SMs
            with contextlib.redirect_stdout(sink):
                cur = snap.project(weight_name="mass", limXY=workloads.LIM_XY, nXY=NXY, limV=workloads.LIM_V, nV=NV,
                                   accumulate=args.accumulate, reducer=reducer, deferred=True)
            if pending[0] is not None:
                finish(pending[0])
            pending[0] = cur
            return None, None
        if reducer is not None and args.reduce == "owner":
            # round-robin owner: step k's grids are reduced to, finalised and smoothed on rank k % N,
            # so that the post-processing load is spread over the ranks instead of piling up on rank 0
            reducer.dst = counter[0] % world
            counter[0] += 1
        with contextlib.redirect_stdout(sink):
            vmap, losvd = snap.project(weight_name="mass", limXY=workloads.LIM_XY, nXY=NXY, limV=workloads.LIM_V,
                                       nV=NV, accumulate=args.accumulate, reducer=reducer, post_stream=post)
            if losvd is not None:
                with (torch.cuda.stream(post) if post is not None else contextlib.nullcontext()):
                    losvd = losvd.convolve_spatial(*SMOOTH_FWHM)
        sink.seek(0)
        sink.truncate()
        return vmap, losvd

    def fence():
        drain()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    # setup, not warm-up: NCCL builds its connections lazily per reduce root, so every owner rank is
    # used once before the W warm-up steps (otherwise a short W leaves connection setup in the timed region)
    for _ in range(world if reducer is not None and args.reduce == "owner" else 0):
        step()
    fence()
    for _ in range(max(3, args.warmup)):
        step()
    fence()
    # the ranks of a sharded run meet at a device-side barrier every step: a cyclic-GC pause of the Python host of any
    # rank (milliseconds with torch loaded) stalls all of them, so the collector rests during the timed region
    import gc
    gc.collect()
    gc.disable()
    sampler = ClockSampler(local) if rank == 0 else None     # rank 0's GPU stands for the node (one line is printed)
    fence()
    engine.KERNEL_TIMER = timer = engine.KernelTimer()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    enqueue_s = [time.perf_counter()]
    for _ in range(args.steps):
        step()
    drain()                                                 # (deferred mode: the last step's finalisation is inside the timed region)
    if sampler is not None:
        sampler.sample_now()
    enqueue_s[0] = time.perf_counter() - enqueue_s[0]       # host time spent launching the K steps (no synchronisation inside)
    if post is not None:
        torch.cuda.current_stream().wait_stream(post)      # the last step's post-processing is inside the timed region
    t1.record()
    fence()
    gc.enable()
    engine.KERNEL_TIMER = None
    clocks = sampler.stop() if sampler is not None else None
    ms = torch.tensor([t0.elapsed_time(t1)], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    total_ms = float(ms.item())
    kernel_ms = timer.total_ms() / max(1, len(timer.pairs))
    kernel_name = timer.kernel or "k_project_bin"
    host_enqueue_ms = enqueue_s[0] * 1e3 / args.steps

    # ---- parity evidence: one more step with int64 fixed-point accumulators through the same
    # kernel and the same reduce; rank 0 then bins the shards of ALL ranks one after the other into one accumulator
    # and compares bits.  At N = 1 the comparison is the snapshot in one launch against two ragged half shards.
    parity = None
    if not args.no_parity:
        parity = parity_check(pnm, engine, workloads, snap, soa, reducer, rank, world, n_local, sink)

    # ---- context (not part of the contract)
    extras = None
    if rank == 0 and n_gpus == 1 and not args.no_extras:
        extras = {}

        def timed(fn, reps):
            for _ in range(3):
                fn()
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(reps):
                fn()
            b.record(); torch.cuda.synchronize()
            return a.elapsed_time(b) / reps

        def plain_step(sn):
            with contextlib.redirect_stdout(sink):
                _, lv = sn.project(weight_name="mass", limXY=workloads.LIM_XY, nXY=NXY, limV=workloads.LIM_V, nV=NV,
                                   accumulate=args.accumulate)
                lv.convolve_spatial(*SMOOTH_FWHM)
            sink.seek(0); sink.truncate()

        # what a user who projects one snapshot once sees: scatter, finalise, transpose, smooth back to back on ONE stream
        ms_np = timed(lambda: plain_step(snap), 10)
        extras["non_pipelined_ms_per_step"] = ms_np
        extras["non_pipelined_value"] = n_local / (ms_np * 1e-3)
        # the same step on a space-filling-curve ordered copy of the snapshot (how tree codes store particles)
        srt = workloads.morton_sort(soa)
        snap_s = pnm.snapshot.from_arrays(srt[:3], srt[3:6], mass=srt[6], output="torch")
        snap_s.rotate(PA=PA, inclination=INCL)
        ms_s = timed(lambda: plain_step(snap_s), 10)
        extras["morton_ordered_snapshot"] = {"ms_per_step": ms_s, "value": n_local / (ms_s * 1e-3), "unit": "particles/s",
                                             "note": "same step, particles pre-sorted along a 3-D Morton curve (sort not timed)"}
        del snap_s, srt

    # ---- end to end: pinned host buffers in, host results out, copies inside the timed region
    e2e = None
    if not args.no_e2e:
        host = torch.empty((7, n_local), dtype=torch.float32, pin_memory=True)
        host.copy_(soa)
        torch.cuda.synchronize()
        rows = [host[k] for k in range(7)]
        out = {"M": torch.empty((NXY, NXY), dtype=torch.float64, pin_memory=True),
               "V": torch.empty((NXY, NXY), dtype=torch.float64, pin_memory=True),
               "S": torch.empty((NXY, NXY), dtype=torch.float64, pin_memory=True),
               "cube": torch.empty((NXY, NXY, NV), dtype=torch.float64, pin_memory=True)}

        def e2e_step():
            return hostpath.project_host(rows, PA, INCL, workloads.LIM_XY, NXY, workloads.LIM_V, NV,
                                         accumulate="f64", smooth_fwhm=SMOOTH_FWHM, reducer=e2e_reducer, out=out)

        e2e_steps = max(1, min(args.steps, args.e2e_steps))
        for _ in range(2):
            e2e_step()
        fence()
        w0 = time.perf_counter()
        for _ in range(e2e_steps):
            e2e_step()
        fence()
        wall = torch.tensor([time.perf_counter() - w0], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(wall, op=dist.ReduceOp.MAX)
        e2e = {"value": n_local * n_gpus * e2e_steps / float(wall.item()), "unit": "particles/s",
               "h2d_bytes_per_step": BYTES_PER_PARTICLE * n_local * n_gpus,
               "d2h_bytes_per_step": 8 * (3 * NXY * NXY + NXY * NXY * NV),
               "steps": e2e_steps, "ms_per_step": float(wall.item()) * 1e3 / e2e_steps,
               "api": "pynmap_b200.hostpath.project_host -> C ABI pnm_project_host" if world == 1 else
                      "pynmap_b200.hostpath.project_host -> pnm_accumulate_host + NCCL reduce",
               "timer": "host perf_counter around synchronous calls, max over ranks"}
        del host

    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))

    # ---- BASELINE.json config 3 at N > 1: 10^9 particles over 8 GPUs (1.25e8 per GPU), 512 x 512 moment maps, NCCL
    # all-reduce of the partial grids; ONE pass, nothing pipelined: the latency a user of velocity_moments() sees
    c3 = None
    if world > 1 and not args.no_extras:
        try:
            c3 = c3_point(pnm, workloads, GridReducer(all_ranks=True), rank, world, sink)
        except torch.cuda.OutOfMemoryError as exc:        # (context only: never take the contract line down)
            c3 = {"error": repr(exc)[:200]}

    # ---- the north star's single-GPU size: 10^9 particles resident on one B200 (28 GB), same grid, same step
    if extras is not None and not args.no_1e9 and n_local < BIG_N:
        del snap, soa
        torch.cuda.empty_cache()
        extras["n1_1e9"] = big_point(pnm, engine, workloads, BIG_N, peak, sink, args.accumulate)

    if rank == 0:
        achieved = BYTES_PER_PARTICLE * n_local / (kernel_ms * 1e-3) / 1e9
        traffic = None
        try:
            traffic = json.load(open(os.path.join(ROOT, "profiles", "roofline_traffic.json"))).get(kernel_name)
        except Exception:
            pass
        line = {
            "metric": METRIC, "value": n_local * n_gpus * args.steps / (total_ms * 1e-3), "unit": "particles/s",
            "n_gpus": n_gpus, "steps": args.steps, "warmup": max(3, args.warmup), "ms_per_step": total_ms / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32",          # particles, rotation and weighted terms in float32 (as in the reference); sums: see config
            "data": "synthetic",
            "config": {"workload": workload_name(n_gpus, n_local), "particles_per_gpu": n_local, "nXY": NXY, "nV": NV,
                       "accumulate": args.accumulate,
                       "accumulators": ("float64; hot pixels / voxels first summed per SM in shared memory as 32-bit fixed-point limbs "
                                        "(every term rounded by < 4.8e-7 of itself), the rest float64 reductions at L2"
                                        if args.accumulate == "f64" else "int64 fixed point (bit-exact, order independent)"),
                       "smoothing_parity": "unpinned (astropy absent: oracle restates convolve() from its documentation)",
                       "parallelism": ("1 GPU" if n_gpus == 1 else
                                       "particle shards x%d + NCCL reduce of partial grids (owner rank = step %% N)" % n_gpus
                                       if args.reduce == "owner" else
                                       "particle shards x%d; the partial grids live in symmetric memory and every rank pulls its "
                                       "nV/N velocity bins of all ranks' grids over NVLink with the copy engines while the next "
                                       "scatter kernel runs (deferred finalisation), sums them in rank order, finalises and smooths "
                                       "them (the smoothed cube stays sharded along V) + all-reduce of the per-pixel map sums "
                                       "(maps on every rank)" % n_gpus
                                       if args.reduce == "peer" else
                                       "particle shards x%d + NCCL reduce-scatter of the partial grids along the velocity axis "
                                       "(every rank finalises and smooths its own nV/N velocity bins; the smoothed cube stays "
                                       "sharded) + all-reduce of the per-pixel map sums (maps on every rank)" % n_gpus),
                       "l2": "inputs (2.8 GB per GPU) are 22x larger than L2, no flush needed",
                       "streams": ("scatter on the main stream; fold / reduce / finalise / smooth on a post-processing "
                                   "stream (overlaps the next step's scatter)" if post is not None else
                                   "one compute stream (scatter k, then finalisation of step k-1) + a copy stream for the "
                                   "peer-to-peer pulls" if deferred else "single stream")},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "traffic_source": "profiles/roofline_traffic.json (ncu --set full of this kernel on "
                         "this workload, static: not measured in this run)" if traffic else None,
                         "kernel": kernel_name, "kernel_ms": kernel_ms, "kernel_launches_timed": len(timer.pairs),
                         "algorithmic_bytes_per_launch": BYTES_PER_PARTICLE * n_local,
                         "peak_source": "MEASURED_PEAKS.json hbm_gbs (of measured)" if peaks else "fallback 6650 GB/s (of fallback)"},
            "cpu_baseline": base,
            "e2e": e2e,
            # our kernels launched by rank 0 inside the timed region: k_project_bin every step (no replica fold: the
            # interleaved cube has a single map replica); k_finalize_vmaps, k_cube_out, k_smooth_axis<0>, <1> on the
            # steps it owns
            "gpu_launches": (5 * args.steps if n_gpus == 1 else
                             args.steps + 4 * len([k for k in range(args.steps) if (k + max(3, args.warmup)) % n_gpus == 0])
                             if args.reduce == "owner" else
                             (7 if args.reduce == "peer" else 6) * args.steps),   # k_project_cube, (k_sum_chunks,) k_slab_sums, k_finalize_sums, k_cube_out_slabs, k_smooth_axis<0>, <1>
            "clocks": clocks,
            "host_enqueue_ms_per_step": host_enqueue_ms,
            "parity": parity,
            "extras": extras if c3 is None else dict(extras or {}, c3=c3),
        }
        print(json.dumps(line), file=json_out, flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "ours", "reference"])
    ap.add_argument("--particles", type=int, default=N_PER_GPU, help="particles per GPU")
    ap.add_argument("--accumulate", default="f64", choices=["f64", "fixed"])
    ap.add_argument("--e2e-steps", type=int, default=5)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-extras", action="store_true")
    ap.add_argument("--no-parity", action="store_true")
    ap.add_argument("--no-1e9", action="store_true", help="skip the 10^9-particle single-GPU point of the extras")
    ap.add_argument("--no-pipeline", action="store_true", help="run post-processing on the scatter stream")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--nccl-max-ctas", type=int, default=4,
                    help="cap on the CTAs of the NCCL reduce, which shares the SMs with the scatter kernel (0 = NCCL's choice)")
    ap.add_argument("--reserve-sms", type=int, default=0,
                    help="SMs the persistent scatter kernel leaves idle for the collective at N > 1")
    ap.add_argument("--reduce", default="peer", choices=["peer", "scatter", "owner"],
                    help="N > 1: peer = copy-engine pulls over NVLink from symmetric memory + local sum (every rank post-processes "
                         "its own velocity bins); scatter = the same with an NCCL reduce-scatter; owner = NCCL reduce to a rotating owner")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    return run_b200(args)


if __name__ == "__main__":
    sys.exit(main())
