"""Device engine behind the pynmap call surface: grids, accumulators and the calls into
libpnm_b200.so.  torch is used for device memory, streams and (in sharded.py) the
process group and symmetric memory -- every arithmetic step of the projection path (rotate, bin, reduce over ranks,
finalise, smooth, and the probes that pick scales / layouts) is a kernel of the library.  What is left to torch: dtype
conversions and copies at ingestion, clearing accumulators, and the one-off per-snapshot flux = mass / ML division.

Layout in HBM
  particles     SoA, one contiguous array per coordinate (x, y, z, vx, vy, vz, w)
  maps acc      [view][iy][ix][4]  8-byte slots {S0 = sum w, S1 = sum w d, S2 = sum w d^2, -}
  cube acc      [view][ix][iy][iv] 8-byte elements (sum w)
  outputs       float64: maps (nY, nX) like grid.py:126-131 after .T, cube (nX, nY, nV) like grid.py:259
"""
import ctypes
import math
import threading
from collections import OrderedDict

import numpy as np
import torch

from . import _lib

_TORCH_TO_PNM = {torch.float32: _lib.F32, torch.float64: _lib.F64}


# ----------------------------------------------------------------------------- small helpers
def is_cuda_tensor(a):
    return isinstance(a, torch.Tensor) and a.is_cuda


def current_device():
    _lib.require_device()
    if not torch.cuda.is_available():
        raise _lib.PnmError("torch sees no CUDA device: pynmap_b200 has no CPU fallback")
    return torch.device("cuda", torch.cuda.current_device())


def stream_ptr():
    """The current CUDA stream of the current device as a void* (the raw query: torch.cuda.current_stream() builds a
    Stream object and costs ~15 us per call, several times per step)."""
    return ctypes.c_void_p(torch._C._cuda_getCurrentRawStream(torch._C._cuda_getDevice()))


def ptr(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else None


def result_dtype(*arrays):
    """np.result_type(float32, inputs...) restricted to the two dtypes the kernels are
    compiled for: the reference's rotation matrix is float32 (rotation.py:107-108), so
    float32 data stay float32 and anything wider (or integer) becomes float64."""
    dts = []
    for a in arrays:
        if a is None:
            continue
        if isinstance(a, torch.Tensor):
            dts.append(torch.empty(0, dtype=a.dtype).numpy().dtype if a.dtype != torch.bfloat16 else np.dtype(np.float32))
        else:
            dts.append(np.asarray(a).dtype)
    dt = np.result_type(np.float32, *dts)
    return torch.float32 if dt == np.float32 else torch.float64


def to_device(a, dtype, device=None):
    """1-D contiguous device tensor of `dtype` from numpy / torch input (copy only if needed)."""
    if a is None:
        return None
    device = device or current_device()
    if isinstance(a, torch.Tensor):
        t = a.detach()
    else:
        arr = np.asarray(a)
        if arr.dtype == np.bool_:
            arr = arr.astype(np.uint8)
        t = torch.from_numpy(np.ascontiguousarray(arr))
    t = t.to(device=device, dtype=dtype, non_blocking=True).contiguous()
    if t.data_ptr() % 32:               # a slice of a larger tensor: the vector-load kernels want aligned arrays
        t = t.clone()
    return t


def alloc_soa(rows, n, dtype, device):
    """(rows, n) device tensor whose rows start on 32-byte boundaries (row pitch rounded up to 32 bytes): the
    vectorised scatter kernels (256-bit / 128-bit loads) want every coordinate array aligned, whatever n is."""
    per = 32 // torch.empty(0, dtype=dtype).element_size()
    pitch = -(-max(int(n), 1) // per) * per
    return torch.empty((rows, pitch), dtype=dtype, device=device)[:, :n]


def to_output(t, want_torch):
    return t if want_torch else t.cpu().numpy()


# ----------------------------------------------------------------------------- bin edges
def expand_nxy(nXY):
    """grid.py:103-107 -- scalar nXY means a square grid; only 1 or 2 entries are legal."""
    if np.size(nXY) == 1:
        n = int(np.asarray(nXY).reshape(-1)[0])
        return n, n
    if np.size(nXY) != 2:
        return None
    return int(nXY[0]), int(nXY[1])


def bin_edges(lo, hi, n):
    """grid.py:119-124: edges padded by half a pixel so that centres are linspace(lo, hi, n).
    Built with the reference's numpy expression on the host: np.linspace edges are not
    exactly uniform, and the kernel compares against these very doubles."""
    half = (hi - lo) / (n - 1) / 2.
    return np.linspace(lo - half, hi + half, n + 1)


class Grid(object):
    """Edge tables for one (limXY, nXY[, limV, nV]) on one device (pnm_grid handle)."""

    def __init__(self, limXY, nx, ny, limV=None, nv=0, device=None):
        self.device = device or current_device()
        self.nx, self.ny, self.nv = int(nx), int(ny), int(nv)
        self.limXY = [float(v) for v in limXY]
        self.limV = None if limV is None else [float(v) for v in limV]
        self.edges_x = bin_edges(limXY[0], limXY[1], self.nx)
        self.edges_y = bin_edges(limXY[2], limXY[3], self.ny)
        self.edges_v = bin_edges(limV[0], limV[1], self.nv) if self.nv else None
        self.handle = ctypes.c_void_p()
        self._mesh_np = self._mesh_dev = self._centres = None
        ex, ey = _lib.doubles(self.edges_x), _lib.doubles(self.edges_y)
        ev = _lib.doubles(self.edges_v) if self.nv else None
        with torch.cuda.device(self.device):
            _lib.call("pnm_grid_create", self.nx, ex, self.ny, ey, self.nv, ev, ctypes.byref(self.handle))

    @property
    def maps_elems(self):
        """8-byte elements of one view's (opaque, replicated) maps accumulator."""
        return int(_lib.load().pnm_maps_acc_elems(self.handle))

    @property
    def map_replicas(self):
        return int(_lib.load().pnm_grid_map_replicas(self.handle))

    @property
    def maps_fold_elems(self):
        """Elements that carry the sums after fold_maps (the first replica)."""
        return self.nx * self.ny * _lib.MAP_SLOTS

    @property
    def cube_elems(self):
        return self.nx * self.ny * self.nv

    @property
    def cube_moments_elems(self):
        """Interleaved cube (CUBE_MOMENTS): records {w even bin, w odd bin, sum w*d, sum w*d*d} per (iv // 2, pixel),
        plus one record slab for the particles outside the V range."""
        return self.nx * self.ny * ((self.nv + 1) // 2 + 1) * 4

    @property
    def record_slabs(self):
        """Record slabs of the interleaved cube: G = ceil(nv / 2) velocity-bin pairs + the out-of-range slab."""
        return (self.nv + 1) // 2 + 1

    @property
    def slab_elems(self):
        return self.nx * self.ny * 4

    def centres(self):
        """Pixel centres (grid.py:141-142, :247-249); fresh arrays, the caller may keep them."""
        if self._centres is None:
            px = np.linspace(self.limXY[0], self.limXY[1], self.nx)
            py = np.linspace(self.limXY[2], self.limXY[3], self.ny)
            pv = np.linspace(self.limV[0], self.limV[1], self.nv) if self.nv else None
            self._centres = (px, py, pv)
        return tuple(None if a is None else a.copy() for a in self._centres)

    def mesh(self, want_torch, share=False):
        """gridX, gridY of the reference (np.meshgrid of the centres, (nY, nX)).  The device copy is
        made once per grid: a per-call pageable upload would stall the stream."""
        if self._mesh_np is None:
            px, py, _ = self.centres()
            self._mesh_np = np.meshgrid(px, py)
        if not want_torch:
            return self._mesh_np[0].copy(), self._mesh_np[1].copy()
        if self._mesh_dev is None:
            self._mesh_dev = tuple(torch.from_numpy(a).to(self.device) for a in self._mesh_np)
        if share:                     # read-only by convention (two clones per call cost two kernel launches)
            return self._mesh_dev
        return self._mesh_dev[0].clone(), self._mesh_dev[1].clone()

    def __del__(self):
        try:
            if self.handle:
                _lib.load().pnm_grid_destroy(self.handle)
        except Exception:
            pass


_GRID_CACHE = OrderedDict()
_GRID_LOCK = threading.Lock()
_GRID_CACHE_MAX = 32


def get_grid(limXY, nx, ny, limV=None, nv=0, device=None):
    device = device or current_device()
    key = (str(device), tuple(float(v) for v in limXY), int(nx), int(ny),
           None if limV is None or not nv else tuple(float(v) for v in limV), int(nv))
    with _GRID_LOCK:
        g = _GRID_CACHE.get(key)
        if g is None:
            g = Grid(limXY, nx, ny, limV if nv else None, nv, device)
            _GRID_CACHE[key] = g
            while len(_GRID_CACHE) > _GRID_CACHE_MAX:
                _GRID_CACHE.popitem(last=False)
        else:
            _GRID_CACHE.move_to_end(key)
    return g


# ----------------------------------------------------------------------------- accumulate mode
class AccMode(object):
    """'f64' (default): float64 accumulators.  'fixed': int64 fixed point, deterministic and bit-exact.
    'f64-exact': float64 accumulators without the shared-memory privatisation of snapshot.project -- every term is
    added with a float64 reduction at L2 (1e-12 relative vs sequential sums; 'f64' rounds the privatised terms to
    2^-25 of the typical term and by less than 4.8e-7 of themselves: each sum within 4.8e-7 x sum|term|)."""

    def __init__(self, name="f64"):
        name = {"float64": "f64", "int64": "fixed", "fixed64": "fixed", "f64_exact": "f64-exact"}.get(str(name), str(name))
        if name not in ("f64", "fixed", "f64-exact"):
            raise ValueError("accumulate must be 'f64', 'f64-exact' or 'fixed', got %r" % (name,))
        self.privatise = name != "f64-exact"
        self.name = name = "f64" if name == "f64-exact" else name
        self.code = _lib.ACC_F64 if name == "f64" else _lib.ACC_FIXED
        self.dtype = torch.float64 if name == "f64" else torch.int64


def fixed_scales(n_total, max_w, max_d):
    """Power-of-two scales for (w, w*d, w*d*d) such that n_total worst-case terms cannot
    overflow int64: n * max|term| * 2^k <= 2^62.  Exactly representable, so `value*scale`
    is exact and only the final rint rounds."""
    out = []
    for bound in (max_w, max_w * max_d, max_w * max_d * max_d):
        bound = max(float(bound), 1e-300) * max(int(n_total), 1)
        k = int(math.floor(62 - math.log2(bound)))
        k = max(-1000, min(k, 1000))
        out.append(math.ldexp(1.0, k))
    return out


def absmax(t):
    """max |t| over finite entries as a python float (one small sync)."""
    if t is None or t.numel() == 0:
        return 0.0
    out = torch.empty(1, dtype=torch.float64, device=t.device)
    _lib.call("pnm_absmax", ptr(t), t.numel(), _TORCH_TO_PNM[t.dtype], ptr(out), stream_ptr())
    return float(out.item())


def order_probe(x, y, z, limit, nsamples=65536):
    """pnm_order_probe: (pairs of consecutive particles closer than ``limit``, pairs examined) on a strided sample --
    one small kernel and one 16-byte read."""
    n = x.numel()
    out = torch.empty(2, dtype=torch.int64, device=x.device)
    stride = max(1, (n - 1) // max(1, nsamples))
    nsamp = min(nsamples, max(0, (n - 1 + stride - 1) // stride))
    _lib.call("pnm_order_probe", ptr(x), ptr(y), ptr(z), n, _TORCH_TO_PNM[x.dtype], stride, nsamp, float(limit), ptr(out),
              stream_ptr())
    close, looked = (int(v) for v in out.tolist())
    return close, looked


# ----------------------------------------------------------------------------- the scatter
def new_accumulators(grid, sums, acc, nviews=1, slab_multiple=1):
    """Zeroed (maps, cube) accumulators.  ``slab_multiple``: the interleaved cube is padded with empty record slabs
    to a multiple of that many slabs (a reduce-scatter over ranks wants equal chunks).  With the interleaved cube (CUBE_MOMENTS) everything lives in the cube
    accumulator and maps is None.  Otherwise, for one view, both live in ONE allocation laid out
    [cube | map replica 0 | map replicas 1..], so that a single memset clears them and, after fold_maps, the leading
    cube + first-replica range is all a multi-GPU reduction has to move (see reduce_span)."""
    want_maps = bool(sums & (_lib.SUM_W | _lib.SUM_WD | _lib.SUM_WD2))
    want_cube = bool(sums & _lib.SUM_CUBE)
    if sums & _lib.CUBE_MOMENTS:
        if slab_multiple > 1 and nviews == 1:
            nslab = -(-grid.record_slabs // slab_multiple) * slab_multiple
            return None, torch.zeros(nslab * grid.slab_elems, dtype=acc.dtype, device=grid.device)
        return None, torch.zeros(nviews * grid.cube_moments_elems, dtype=acc.dtype, device=grid.device)
    if nviews == 1 and want_maps and want_cube:
        ncube, nmaps = grid.cube_elems, grid.maps_elems
        pad = ncube % 2                  # keep the maps part 16-byte aligned (TMA bulk reductions)
        flat = torch.zeros(ncube + pad + nmaps, dtype=acc.dtype, device=grid.device)
        return flat[ncube + pad:], flat[:ncube]
    maps = cube = None
    if want_maps:
        maps = torch.zeros(nviews * grid.maps_elems, dtype=acc.dtype, device=grid.device)
    if want_cube:
        cube = torch.zeros(nviews * grid.cube_elems, dtype=acc.dtype, device=grid.device)
    return maps, cube


def reduce_span(grid, maps, cube):
    """The tensors a sum-reduction over ranks must cover once fold_maps has run: one contiguous
    view when (maps, cube) came from a joint allocation, else the two pieces."""
    folded = maps[:grid.maps_fold_elems] if maps is not None else None
    if maps is not None and cube is not None and maps.untyped_storage().data_ptr() == cube.untyped_storage().data_ptr():
        gap = maps.storage_offset() - (cube.storage_offset() + cube.numel())
        if 0 <= gap <= 1:
            joint = torch.as_strided(cube, (cube.numel() + gap + grid.maps_fold_elems,), (1,), cube.storage_offset())
            return (joint,)
    return tuple(t for t in (folded, cube) if t is not None)


def _scales_arg(acc, scales):
    if acc.code == _lib.ACC_FIXED:
        if scales is None:
            raise ValueError("fixed-point accumulation needs scales")
        return _lib.doubles(scales)
    return None


class KernelTimer(object):
    """Optional CUDA-event bracket around the scatter kernel (bench.py's roofline figure).
    Events are recorded on the launching stream; read with ``total_ms`` after a synchronise."""

    def __init__(self):
        self.pairs = []
        self.kernel = None          # name of the kernel the last bracket timed

    def bracket(self, kernel="k_project_bin"):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        self.pairs.append((a, b))
        self.kernel = kernel
        return a, b

    def total_ms(self):
        return sum(a.elapsed_time(b) for a, b in self.pairs)


KERNEL_TIMER = None      # set to a KernelTimer to time every pnm_project_bin launch


def project_bin(grid, soa, w, mats, nchain, nviews, sums, acc, scales, maps, cube, mask=None, mask_mode=0):
    """pnm_project_bin: soa = (x, y, z, vx, vy, vz) device tensors of one dtype."""
    x = soa[0]
    mats = np.ascontiguousarray(np.asarray(mats, dtype=np.float32).reshape(-1))
    assert mats.size == 9 * nchain * nviews
    if KERNEL_TIMER is not None:
        t0, t1 = KERNEL_TIMER.bracket()
        t0.record()
        try:
            _project_bin_call(grid, x, soa, w, mats, nchain, nviews, sums, acc, scales, maps, cube, mask, mask_mode)
        finally:
            t1.record()
        return
    _project_bin_call(grid, x, soa, w, mats, nchain, nviews, sums, acc, scales, maps, cube, mask, mask_mode)


def _project_bin_call(grid, x, soa, w, mats, nchain, nviews, sums, acc, scales, maps, cube, mask, mask_mode):
    _lib.call("pnm_project_bin", grid.handle, x.numel(), _TORCH_TO_PNM[x.dtype],
              ptr(soa[0]), ptr(soa[1]), ptr(soa[2]), ptr(soa[3]), ptr(soa[4]), ptr(soa[5]),
              ptr(w), ptr(mask), mask_mode,
              mats.ctypes.data_as(ctypes.POINTER(ctypes.c_float)), nchain, nviews,
              sums, acc.code, _scales_arg(acc, scales), ptr(maps), ptr(cube), stream_ptr())


class CubePlan(object):
    """pnm_cube_plan handle: which pixels' moments and which voxels' weights k_project_cube accumulates in shared
    memory first (csrc/pnm_cube.cuh, csrc/pnm_plan.cuh).  Built on the device from a sample of the particles; it
    changes the speed of project_cube, never its result."""

    def __init__(self, grid):
        self.grid = grid
        self.handle = ctypes.c_void_p()
        with torch.cuda.device(grid.device):
            _lib.call("pnm_cube_plan_create", grid.handle, ctypes.byref(self.handle))

    def build(self, soa, w, mat):
        m = np.ascontiguousarray(np.asarray(mat, dtype=np.float32).reshape(9))
        _lib.call("pnm_cube_plan_build", self.grid.handle, self.handle, soa[0].numel(),
                  ptr(soa[0]), ptr(soa[1]), ptr(soa[2]), ptr(soa[3]), ptr(soa[4]), ptr(soa[5]), ptr(w),
                  m.ctypes.data_as(ctypes.POINTER(ctypes.c_float)), stream_ptr())
        self._info = None
        return self

    def reserve_sms(self, nsm):
        """Leave ``nsm`` SMs idle in project_cube launches with this plan (room for a collective of another stream)."""
        if nsm != getattr(self, "_reserved", 0):
            _lib.call("pnm_cube_plan_set_reserved_sms", self.handle, int(nsm))
            self._reserved = int(nsm)
        return self

    def info(self):
        """dict(bx0, by0, bw, bh, nslots, level, captured, sampled, mean_abs=(|w|, |w d|, |w d d|), smem_scales) of
        the last build -- one stream synchronisation, then cached."""
        if getattr(self, "_info", None) is None:
            out, stats = (ctypes.c_int32 * 8)(), (ctypes.c_double * 6)()
            _lib.call("pnm_cube_plan_info", self.handle, out, stats, stream_ptr())
            self._info = dict(zip(("bx0", "by0", "bw", "bh", "nslots", "level", "captured", "sampled"), [int(v) for v in out]))
            self._info["mean_abs"] = tuple(float(v) for v in stats[:3])
            self._info["smem_scales"] = tuple(float(v) for v in stats[3:])
        return self._info

    def __del__(self):
        try:
            if self.handle:
                _lib.load().pnm_cube_plan_destroy(self.handle)
        except Exception:
            pass


CUBE_TYP_BITS = 24            # csrc/pnm_cube.cuh kCubeTypBits


def cube_fixed_scales(n_total, max_w, max_d, mean_abs):
    """int64 fixed-point scales for runs through k_project_cube: engine.fixed_scales (no overflow for n_total
    worst-case terms), capped so that the typical term (the plan's sample mean of |term|) lands in [2^24, 2^25) --
    the kernel keeps a term in shared memory only when its scaled value fits 32 bits.  The cap trades resolution
    (2^-25 of the typical term instead of ~2^-36) for speed; any power of two gives bit-exact, order-independent sums."""
    out = fixed_scales(n_total, max_w, max_d)
    for k, m in enumerate(mean_abs):
        if m > 0.0 and math.isfinite(m):
            out[k] = min(out[k], math.ldexp(1.0, CUBE_TYP_BITS - int(math.floor(math.log2(m)))))
    return [min(max(s, 2.0 ** -120), 2.0 ** 120) for s in out]


def cube_scales_ok(scales):
    """True when int64 fixed-point scales can be applied in float32 (project_cube), i.e. 2^-120 <= 2^k <= 2^120."""
    return scales is None or all(2.0 ** -120 <= float(s) <= 2.0 ** 120 for s in scales)


def cube_path_ok(grid, soa, w):
    """k_project_cube takes float32 particles whose arrays are 16-byte aligned, and a grid with a V axis."""
    if grid.nv < 1 or soa[0].dtype != torch.float32 or grid.cube_moments_elems >= 2 ** 31 or soa[0].numel() >= 2 ** 32:
        return False
    return all(t is None or t.data_ptr() % 16 == 0 for t in tuple(soa) + (w,))


def project_cube(grid, plan, soa, w, mat, acc, scales, cube):
    """pnm_project_cube: float32 SoA particles, one view; adds into the interleaved cube accumulator ``cube``
    (grid.cube_moments_elems elements of acc.dtype).  Launches that share a plan must share a stream (the plan
    owns the kernel's work counter)."""
    x = soa[0]
    if x.dtype != torch.float32:
        raise TypeError("project_cube takes float32 particles")
    m = np.ascontiguousarray(np.asarray(mat, dtype=np.float32).reshape(9))
    args = (grid.handle, plan.handle, x.numel(), ptr(soa[0]), ptr(soa[1]), ptr(soa[2]), ptr(soa[3]), ptr(soa[4]),
            ptr(soa[5]), ptr(w), m.ctypes.data_as(ctypes.POINTER(ctypes.c_float)), acc.code,
            _scales_arg(acc, scales), ptr(cube), stream_ptr())
    if KERNEL_TIMER is not None:
        t0, t1 = KERNEL_TIMER.bracket("k_project_cube")
        t0.record()
        try:
            _lib.call("pnm_project_cube", *args)
        finally:
            t1.record()
        return
    _lib.call("pnm_project_cube", *args)


def bin_points(grid, x, y, d, w, sums, acc, scales, maps, cube, mask=None, mask_mode=0):
    """pnm_bin_points: already projected coordinates."""
    _lib.call("pnm_bin_points", grid.handle, x.numel(), _TORCH_TO_PNM[x.dtype],
              ptr(x), ptr(y), ptr(d), ptr(w), ptr(mask), mask_mode,
              sums, acc.code, _scales_arg(acc, scales), ptr(maps), ptr(cube), stream_ptr())


def finalize_vmaps(grid, maps, cube, sums, acc, scales, want_raw=False):
    """-> M, V, S (and S1, S2 when want_raw) as float64 (nY, nX) device tensors."""
    shape = (grid.ny, grid.nx)
    M, V, S = (torch.empty(shape, dtype=torch.float64, device=grid.device) for _ in range(3))
    S1 = S2 = None
    if want_raw:
        S1, S2 = (torch.empty(shape, dtype=torch.float64, device=grid.device) for _ in range(2))
    _lib.call("pnm_finalize_vmaps", grid.handle, ptr(maps), ptr(cube), sums, acc.code, _scales_arg(acc, scales),
              ptr(M), ptr(V), ptr(S), ptr(S1), ptr(S2), stream_ptr())
    return (M, V, S, S1, S2) if want_raw else (M, V, S)


def finalize_map(grid, maps, acc, scales):
    shape = (grid.ny, grid.nx)
    W, D, DW = (torch.empty(shape, dtype=torch.float64, device=grid.device) for _ in range(3))
    _lib.call("pnm_finalize_map", grid.handle, ptr(maps), acc.code, _scales_arg(acc, scales),
              ptr(W), ptr(D), ptr(DW), stream_ptr())
    return W, D, DW


def finalize_cube(grid, cube, acc, scales, sums=0):
    """cube accumulator (velocity-major, or interleaved with the moments when ``sums`` has CUBE_MOMENTS) ->
    float64 (nX, nY, nV) device tensor (np.histogramdd layout)."""
    out = torch.empty((grid.nx, grid.ny, grid.nv), dtype=torch.float64, device=grid.device)
    scale = float(scales[0]) if acc.code == _lib.ACC_FIXED else 1.0
    _lib.call("pnm_finalize_cube", grid.handle, ptr(cube), int(sums), acc.code, scale, ptr(out), stream_ptr())
    return out


def sum_chunks(chunks, nchunks, per, acc):
    """pnm_sum_chunks: ``chunks`` holds nchunks arrays of ``per`` accumulator elements back to back -> their sum, added
    in index order."""
    out = torch.empty(per, dtype=acc.dtype, device=chunks.device)
    _lib.call("pnm_sum_chunks", ptr(chunks), int(nchunks), int(per), acc.code, ptr(out), stream_ptr())
    return out


def sum_peers(ptrs, n, acc, device):
    """pnm_sum_peers: arrays of ``n`` accumulator elements at the device addresses ``ptrs`` (this device's or peers')
    -> their sum, added in list order."""
    out = torch.empty(n, dtype=acc.dtype, device=device)
    arr = (ctypes.c_void_p * len(ptrs))(*ptrs)
    _lib.call("pnm_sum_peers", arr, len(ptrs), int(n), acc.code, ptr(out), stream_ptr())
    return out


def slab_sums(grid, slabs, g0, g1, acc, out=None):
    """pnm_slab_sums: this rank's share of (sum w, sum w*d, sum w*d*d) per pixel from the record slabs [g0, g1) it
    holds (``slabs`` starts at slab g0) -> (3, nx*ny) tensor of the accumulator type."""
    if out is None:
        out = torch.empty((3, grid.nx * grid.ny), dtype=acc.dtype, device=grid.device)
    _lib.call("pnm_slab_sums", grid.handle, ptr(slabs), int(g0), int(g1), acc.code, ptr(out), stream_ptr())
    return out


def finalize_sums(grid, sums3, acc, scales):
    """pnm_finalize_sums: total (sum w, sum w*d, sum w*d*d) per pixel -> M, V, S as float64 (nY, nX) tensors."""
    shape = (grid.ny, grid.nx)
    M, V, S = (torch.empty(shape, dtype=torch.float64, device=grid.device) for _ in range(3))
    _lib.call("pnm_finalize_sums", grid.handle, ptr(sums3), acc.code, _scales_arg(acc, scales), ptr(M), ptr(V), ptr(S),
              stream_ptr())
    return M, V, S


def slab_vrange(grid, g0, g1):
    """Velocity bins [iv0, iv1) carried by the record slabs [g0, g1)."""
    return min(2 * g0, grid.nv), min(2 * g1, grid.nv)


def finalize_cube_slabs(grid, slabs, g0, g1, acc, scales):
    """pnm_finalize_cube_slabs: the velocity bins of the record slabs [g0, g1) -> float64 (nX, nY, bins) tensor
    (possibly with zero bins: a rank that only holds padding or the out-of-range slab)."""
    iv0, iv1 = slab_vrange(grid, g0, g1)
    out = torch.empty((grid.nx, grid.ny, iv1 - iv0), dtype=torch.float64, device=grid.device)
    scale = float(scales[0]) if acc.code == _lib.ACC_FIXED else 1.0
    if iv1 > iv0:
        _lib.call("pnm_finalize_cube_slabs", grid.handle, ptr(slabs), int(g0), int(g1), acc.code, scale, ptr(out), stream_ptr())
    return out


def fold_maps(grid, maps, acc, clear=True):
    """Sum the map replicas into the first one; returns the view that carries the totals (what a
    multi-GPU reduction has to move).  ``clear=False`` leaves the other replicas untouched: cheaper,
    but the buffer must then be finalised with ``sums | MAPS_FOLDED`` and not accumulated into again."""
    _lib.call("pnm_fold_maps", grid.handle, ptr(maps), acc.code, 1 if clear else 0, stream_ptr())
    return maps[:grid.maps_fold_elems]


# ----------------------------------------------------------------------------- smoothing
FWHM_TO_SIGMA = 1.0 / (2.0 * math.sqrt(2.0 * math.log(2.0)))   # astropy.stats.gaussian_fwhm_to_sigma


def gaussian_support(sig_a, sig_b):
    """astropy Gaussian2DKernel default size: ceil(8 * max sigma) bumped to odd."""
    s = int(math.ceil(8 * max(sig_a, sig_b)))
    return s + 1 - s % 2


_TAPS_CACHE = {}


def gaussian_taps(sigma_pix, support):
    """Normalised 1-D factor of the (separable) truncated 2-D Gaussian, float64 on the host (cached: a pipeline
    smooths every step with the same widths)."""
    key = (float(sigma_pix), int(support))
    g = _TAPS_CACHE.get(key)
    if g is None:
        half = (support - 1) // 2
        i = np.arange(-half, half + 1, dtype=np.float64)
        with np.errstate(divide="ignore", invalid="ignore"):
            g = np.exp(-0.5 * (i / sigma_pix) ** 2)
            g = g / g.sum()
        if len(_TAPS_CACHE) > 64:
            _TAPS_CACHE.clear()
        _TAPS_CACHE[key] = g
    return g.copy()


def smooth_cube(cube, taps_axis0, taps_axis1):
    """pnm_smooth_cube on a float64 (nX, nY, nV) device tensor -> new tensor."""
    cube = cube.contiguous()
    nx, ny, nv = cube.shape
    out = torch.empty_like(cube)
    if cube.numel() == 0:                # (a rank of a slab-sharded run may hold no velocity bins)
        return out
    work = torch.empty_like(cube)
    t0, t1 = _lib.doubles(taps_axis0), _lib.doubles(taps_axis1)
    _lib.call("pnm_smooth_cube", ptr(cube), ptr(out), ptr(work), nx, ny, nv, t0, len(taps_axis0), t1,
              len(taps_axis1), stream_ptr())
    return out


def rebin_cube(cube, factorx, factory, mean=False):
    """pnm_rebin_cube on a float64 (nX, nY, nV) device tensor -> (nX//fx, nY//fy, nV)."""
    cube = cube.contiguous()
    nx, ny, nv = cube.shape
    out = torch.empty((nx // factorx, ny // factory, nv), dtype=torch.float64, device=cube.device)
    _lib.call("pnm_rebin_cube", ptr(cube), ptr(out), nx, ny, nv, int(factorx), int(factory), 1 if mean else 0,
              stream_ptr())
    return out


def cube_moments(cube, binv):
    cube = cube.contiguous()
    nx, ny, nv = cube.shape
    binv = np.asarray(binv, dtype=np.float64)
    bv = to_device(binv, torch.float64, cube.device)
    empty_m2 = float(np.min(np.diff(binv)) / 3.0)        # misc_functions.py:81-84
    m0, m1, m2 = (torch.empty((nx, ny), dtype=torch.float64, device=cube.device) for _ in range(3))
    _lib.call("pnm_cube_moments", ptr(cube), nx, ny, nv, ptr(bv), empty_m2, ptr(m0), ptr(m1), ptr(m2), stream_ptr())
    return m0, m1, m2


def fit_gh_defaults(binv, deg_gh):
    """Start values (NaN = from the profile's moments), lower and upper limits of fit_losvd.py:157-163."""
    binv = np.asarray(binv, dtype=np.float64).ravel()
    start = np.full(deg_gh, 0.02)
    start[:2] = np.nan
    lo = np.full(deg_gh, -0.2)
    lo[:2] = (np.min(binv), np.min(np.diff(binv)) / 3.0)
    hi = np.full(deg_gh, 0.2)
    hi[:2] = (np.max(binv), (np.max(binv) - np.min(binv)) / 3.0)
    return start, lo, hi


def fit_gh(profiles, binv, deg_gh=4, err=None, start=None, lo=None, hi=None, fixed=None,
           ftol=1e-10, xtol=1e-10, gtol=1e-10, maxiter=200, want_bestfit=True):
    """pnm_fit_gh on ``profiles`` (..., nv): Gauss-Hermite fit of every profile (row f4;
    fit_losvd.py:84-235).  ``start`` / ``lo`` / ``hi`` / ``fixed`` are per-parameter arrays of
    deg_gh entries (defaults: fit_gh_defaults, nothing fixed).  Returns (gh (deg_gh+1, ...),
    bestfit (..., nv) or None, status, niter, chi2) as device tensors."""
    profiles = profiles.contiguous()
    lead, nv = tuple(profiles.shape[:-1]), int(profiles.shape[-1])
    nspax = int(np.prod(lead)) if lead else 1
    binv = np.asarray(binv, dtype=np.float64).ravel()
    if binv.size != nv:
        raise ValueError("fit_gh: the velocity axis has {} points, the profiles {}".format(binv.size, nv))
    dev = profiles.device
    bv = to_device(binv, torch.float64, dev)
    d_start, d_lo, d_hi = fit_gh_defaults(binv, deg_gh)
    arrs = []
    for given, default in ((start, d_start), (lo, d_lo), (hi, d_hi)):
        a = np.ascontiguousarray(default if given is None else given, dtype=np.float64)
        if a.shape != (deg_gh,):
            raise ValueError("fit_gh: per-parameter arrays need deg_gh = {} entries".format(deg_gh))
        arrs.append(a)
    fx = np.ascontiguousarray(np.zeros(deg_gh) if fixed is None else fixed).astype(np.int32)
    if fx.shape != (deg_gh,):
        raise ValueError("fit_gh: per-parameter arrays need deg_gh = {} entries".format(deg_gh))
    if err is not None:
        err = err.contiguous()
        if tuple(err.shape) != tuple(profiles.shape):
            raise ValueError("fit_gh: err must have the shape of the profiles")
    gh = torch.empty((deg_gh + 1,) + lead, dtype=torch.float64, device=dev)
    best = torch.empty(lead + (nv,), dtype=torch.float64, device=dev) if want_bestfit else None
    status = torch.empty(lead, dtype=torch.int32, device=dev)
    niter = torch.empty(lead, dtype=torch.int32, device=dev)
    chi2 = torch.empty(lead, dtype=torch.float64, device=dev)
    dptr = lambda a: a.ctypes.data_as(ctypes.POINTER(ctypes.c_double))  # noqa: E731
    _lib.call("pnm_fit_gh", ptr(profiles), ptr(err) if err is not None else None, nspax, nv, ptr(bv), int(deg_gh),
              dptr(arrs[0]), dptr(arrs[1]), dptr(arrs[2]), fx.ctypes.data_as(ctypes.POINTER(ctypes.c_int32)),
              0.02, float(np.min(np.diff(binv)) / 3.0), float(ftol), float(xtol), float(gtol), int(maxiter),
              ptr(gh), ptr(best) if best is not None else None, ptr(status), ptr(niter), ptr(chi2), stream_ptr())
    return gh, best, status, niter, chi2


def rotate_soa(soa3, mat, out=None):
    """pnm_rotate on three SoA tensors -> three new tensors (rows of a (3, N) tensor with aligned rows)."""
    x, y, z = soa3
    if out is None:
        out = alloc_soa(3, x.numel(), x.dtype, x.device)
    m = np.ascontiguousarray(np.asarray(mat, dtype=np.float32).reshape(9))
    _lib.call("pnm_rotate", ptr(x), ptr(y), ptr(z), ptr(out[0]), ptr(out[1]), ptr(out[2]), x.numel(),
              _TORCH_TO_PNM[x.dtype], m.ctypes.data_as(ctypes.POINTER(ctypes.c_float)), stream_ptr())
    return out
