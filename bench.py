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
import threading
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
    line = {
        "impl": "reference", "metric": METRIC, "value": base["value"], "unit": "particles/s",
        "n_gpus": args.gpus, "steps": steps, "warmup": warmup, "ms_per_step": t_step * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_name(args.gpus, args.particles), "timed_steps": steps_run,
                   "note": "CPU port of the reference path (pure-Python reference: numpy calls restated in oracle/)"},
        "cpu_baseline": base,
        "e2e": {"value": base["value"], "unit": "particles/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))
    return 0


# ------------------------------------------------------------------------------------------ clocks
class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons through NVML while the timed region runs."""
    REASONS = {0x4: "sw_power_cap", 0x8: "hw_slowdown", 0x20: "sw_thermal_slowdown", 0x40: "hw_thermal_slowdown",
               0x80: "hw_power_brake_slowdown", 0x2: "applications_clocks_setting", 0x10: "sync_boost"}

    def __init__(self, index, period=0.005):
        super().__init__(daemon=True)
        self.index, self.period = index, period
        self.samples, self.mask, self.max_mhz, self.ok = [], 0, None, False
        self._stop_evt = threading.Event()
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception:
            self.ok = False

    def run(self):
        if not self.ok:
            return
        get_reasons = getattr(self.nv, "nvmlDeviceGetCurrentClocksEventReasons", None) or \
            getattr(self.nv, "nvmlDeviceGetCurrentClocksThrottleReasons")
        while not self._stop_evt.is_set():
            try:
                self.samples.append(self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM))
                self.mask |= int(get_reasons(self.h))
            except Exception:
                pass
            time.sleep(self.period)

    def stop(self):
        self._stop_evt.set()
        self.join(timeout=2)
        reasons = [name for bit, name in self.REASONS.items() if self.mask & bit]
        return {"sm_mhz": float(np.median(self.samples)) if self.samples else None, "sm_max_mhz": self.max_mhz,
                "reasons": reasons, "samples": len(self.samples)}


# ------------------------------------------------------------------------------------------ our arm
def run_b200(args):
    import torch
    import torch.distributed as dist

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
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
    from pynmap_b200.sharded import GridReducer

    pnm._lib.require_device()
    soa = workloads.make_snapshot(n_local, bulge_fraction=0.2, seed=SEED + 1000 * rank, device="cuda")
    snap = pnm.snapshot.from_arrays(soa[:3], soa[3:6], mass=soa[6], output="torch")
    snap.rotate(PA=PA, inclination=INCL)
    reducer = GridReducer() if world > 1 else None
    sink = io.StringIO()

    # post-processing stream: replica fold, NCCL reduce, finalisation and smoothing of step k run
    # next to the scatter kernel of step k+1 (all K steps are complete before the timer stops)
    post = None if args.no_pipeline else torch.cuda.Stream(priority=-1)

    counter = [0]

    def step():
        if reducer is not None:
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
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    # setup, not warm-up: NCCL builds its connections lazily per reduce root, so every owner rank is
    # used once before the W warm-up steps (otherwise a short W leaves connection setup in the timed region)
    for _ in range(world if reducer is not None else 0):
        step()
    fence()
    for _ in range(max(3, args.warmup)):
        step()
    fence()
    sampler = ClockSampler(local)
    sampler.start()
    engine.KERNEL_TIMER = timer = engine.KernelTimer()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    for _ in range(args.steps):
        step()
    if post is not None:
        torch.cuda.current_stream().wait_stream(post)      # the last step's post-processing is inside the timed region
    t1.record()
    fence()
    engine.KERNEL_TIMER = None
    clocks = sampler.stop()
    ms = torch.tensor([t0.elapsed_time(t1)], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    total_ms = float(ms.item())
    kernel_ms = timer.total_ms() / max(1, len(timer.pairs))

    # ---- context (not part of the contract): the same step on a space-filling-curve ordered copy of the
    # snapshot (how tree codes store particles); consecutive particles then share pixels and merge in registers
    extras = None
    if rank == 0 and n_gpus == 1 and not args.no_extras:
        srt = workloads.morton_sort(soa)
        snap_s = pnm.snapshot.from_arrays(srt[:3], srt[3:6], mass=srt[6], output="torch")
        snap_s.rotate(PA=PA, inclination=INCL)
        def step_sorted():
            with contextlib.redirect_stdout(sink):
                _, lv = snap_s.project(weight_name="mass", limXY=workloads.LIM_XY, nXY=NXY, limV=workloads.LIM_V, nV=NV,
                                       accumulate=args.accumulate)
                lv.convolve_spatial(*SMOOTH_FWHM)
            sink.seek(0); sink.truncate()
        for _ in range(3):
            step_sorted()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(10):
            step_sorted()
        b.record(); torch.cuda.synchronize()
        ms_s = a.elapsed_time(b) / 10
        extras = {"morton_ordered_snapshot": {"ms_per_step": ms_s, "value": n_local / (ms_s * 1e-3), "unit": "particles/s",
                                              "note": "same step, particles pre-sorted along a 3-D Morton curve (sort not timed)"}}
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
                                         accumulate="f64", smooth_fwhm=SMOOTH_FWHM, reducer=reducer, out=out)

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

    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        peak = float(peaks.get("hbm_gbs", 6650.0))
        achieved = BYTES_PER_PARTICLE * n_local / (kernel_ms * 1e-3) / 1e9
        traffic = None
        try:
            traffic = json.load(open(os.path.join(ROOT, "profiles", "roofline_traffic.json"))).get("k_project_bin")
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
                       "accumulators": "float64 (L2 reductions)" if args.accumulate == "f64" else "int64 fixed point (bit-exact, order independent)",
                       "parallelism": "particle shards x%d + NCCL reduce of partial grids (owner rank = step %% N)" % n_gpus
                       if n_gpus > 1 else "1 GPU",
                       "l2": "inputs (2.8 GB per GPU) are 22x larger than L2, no flush needed",
                       "streams": "scatter on the main stream; fold / reduce / finalise / smooth on a post-processing "
                                  "stream (overlaps the next step's scatter)" if post is not None else "single stream"},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "kernel": "k_project_bin", "kernel_ms": kernel_ms,
                         "algorithmic_bytes_per_launch": BYTES_PER_PARTICLE * n_local,
                         "peak_source": "MEASURED_PEAKS.json hbm_gbs (of measured)" if peaks else "fallback 6650 GB/s (of fallback)",
                         # the unit that actually bounds the scatter: L2 takes one request per
                         # distinct 32-byte sector of a reduction instruction, ~1.87e11 scattered requests/s chip-wide
                         # (scripts/red_pair_bench.cu, profiles/r01_micro_tma_mix.log); with the interleaved cube all
                         # three sums of a particle share a sector: one request per particle
                         "l2_reduction_requests": {"per_particle": 1.0, "achieved_per_s": 1.0 * n_local / (kernel_ms * 1e-3),
                                                   "peak_per_s": 1.87e11, "frac": 1.0 * n_local / (kernel_ms * 1e-3) / 1.87e11}},
            "cpu_baseline": base,
            "e2e": e2e,
            # our kernels launched by rank 0 inside the timed region: k_project_bin every step (no replica fold: the
            # interleaved cube has a single map replica); k_finalize_vmaps, k_cube_out, k_smooth_axis<0>, <1> on the
            # steps it owns
            "gpu_launches": (5 * args.steps if n_gpus == 1 else
                             args.steps + 4 * len([k for k in range(args.steps) if (k + max(3, args.warmup)) % n_gpus == 0])),
            "clocks": clocks,
            "extras": extras,
        }
        print(json.dumps(line))
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
    ap.add_argument("--no-pipeline", action="store_true", help="run post-processing on the scatter stream")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--nccl-max-ctas", type=int, default=4,
                    help="cap on the CTAs of the NCCL reduce, which shares the SMs with the scatter kernel (0 = NCCL's choice)")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    return run_b200(args)


if __name__ == "__main__":
    sys.exit(main())
