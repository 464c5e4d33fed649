"""User-facing object model with the reference's interface (pynmap/pynmap.py:40-474):
``snapshot`` (read / reset / rotate / velocity_moments / create_map / create_losvds) and the
``SnapMap`` result container.

B200-native differences that do not change the call surface:
  * particles live in HBM as SoA tensors (x, y, z, vx, vy, vz, mass/flux);
  * ``rotate`` only records the float32 matrix; ``velocity_moments`` / ``create_map`` /
    ``create_losvds`` run ONE fused rotate+project+bin kernel over the original arrays
    (each compounded rotation is re-applied per particle in registers with the same
    rounding as a materialised rotation, so the results are bit-identical);
  * ``pos`` / ``vel`` are materialised lazily (pnm_rotate) only if the user reads them.
Additions: ``snapshot.from_arrays`` (tensors instead of an ascii file), the ``accumulate``
keyword ('f64' | 'fixed'), ``project`` (maps + LOSVD cube from a single read) and
``project_views`` (many viewing angles from a single read).
"""
import os
from os.path import join as joinpath

import numpy as np
import torch

from . import _lib, engine
from .grid import point_2vprofiles, points_2losvds, points_2map, points_2vmaps
from .losvd import LOSVD
from .rotation import rotation_matrix

# Default scales and limits (pynmap/pynmap.py:31-37)
default_scale = 1000.0
default_Vscale = 1000.0
default_limXY = [-default_scale, default_scale, -default_scale, default_scale]
default_limV = [-default_Vscale, default_Vscale]
default_npoint = 100
default_nVpoint = 100


def convert_to_list(a):
    """pynmap/misc_io.py:359-368: a scalar angle becomes ``[np.float32(a)]`` -- which is why
    the rotation trigonometry of snapshot.rotate runs in float32."""
    return [np.float32(a)]


def read_ascii_files(filename):
    """One particle per line: x y z Vx Vy Vz [mass age Z/H ...] (pynmap/misc_io.py:143-182)."""
    if os.path.isfile(filename):
        try:
            data = np.loadtxt(filename).T
            print("Found {} columns in file".format(data.shape[0]))
            pos = np.vstack(data[:3]).T
            vel = np.vstack(data[3:6]).T
            return pos, vel, data[6:]
        except ValueError:
            print("ERROR: could not read content of the input file")
    else:
        print("Input file not found: {0}".format(filename))
    return None, None, None


dic_io = {"Magneticum": read_ascii_files, "ascii": read_ascii_files}


class SnapMap(object):
    """Result container (pynmap/pynmap.py:40-63): attributes come from ``input_dic``."""

    def __init__(self, name="", info=None, input_dic={}):
        self._info = info
        self.name = name
        for item in input_dic.keys():
            setattr(self, item, input_dic[item])
        self.input_keys = list(input_dic.keys())

    def info(self):
        print("Name {0}: {1}".format(self.name, self._info))
        print("List of keys: {}".format(self.input_keys))

    @property
    def extent(self):
        if hasattr(self, 'X') & hasattr(self, 'Y'):
            X, Y = self.X, self.Y
            if isinstance(X, torch.Tensor):
                return [float(X.min()), float(X.max()), float(Y.min()), float(Y.max())]
            return [np.min(X), np.max(X), np.min(Y), np.max(Y)]
        return None


class SnapMapList(list):
    """A list of SnapMap with a free-text ``info`` (pynmap/pynmap.py:65-96): ``SnapMapList(maps, info=...)`` or
    ``SnapMapList(map1, map2, ...)``; calling the list with keywords sets attributes."""

    def __init__(self, *args, **kwargs):
        items = args[0] if len(args) == 1 and hasattr(args[0], '__iter__') else args
        super(SnapMapList, self).__init__(items)
        self._info = ""
        self(**kwargs)

    def __call__(self, **kwargs):
        if 'info' in kwargs:
            self._info = kwargs.pop('info')
        self.__dict__.update(kwargs)
        return self

    def update(self, **kwargs):
        return self(**kwargs)

    def info(self):
        print(self._info)

    def print(self):
        for item in self:
            print(item.name)


class snapmap(object):
    """A named 2-D array that can be written to / read from an ascii file (pynmap/pynmap.py:477-504)."""

    def __init__(self, name="dummy", data=np.zeros((2, 2))):
        self.name = name
        self.data = data

    def write_ascii(self, filename):
        np.savetxt(filename, _to_numpy(self.data))

    def read_ascii(self, filename):
        self.data = np.loadtxt(filename)

    def write_pickle(self, filename):
        import pickle
        with open(filename, 'wb') as f:
            pickle.dump(_to_numpy(self.data), f)

    def read_pickle(self, filename):
        import pickle
        with open(filename, 'rb') as f:
            self.data = pickle.load(f)


def _to_numpy(a):
    return a.detach().cpu().numpy() if isinstance(a, torch.Tensor) else np.asarray(a)


class PendingProjection(object):
    """What snapshot.project(deferred=True) returns: the scatter kernel has been launched; ``result()`` launches the
    reduction over ranks (if any) and the finalisation, once, and returns (SnapMap, LOSVD)."""

    def __init__(self, snap, grid, maps, cube, sums, acc, scales, reducer, token):
        self._args = (snap, grid, maps, cube, sums, acc, scales, reducer, token)
        self._out = None

    def result(self):
        if self._args is not None:
            snap, grid, maps, cube, sums, acc, scales, reducer, token = self._args
            self._args = None
            if token is not None:
                self._out = snap._project_post_slabs(grid, cube, acc, scales, reducer, token)
            else:
                self._out = snap._project_post(grid, maps, cube, sums, acc, scales, reducer)
        return self._out


class snapshot(object):
    """N-body snapshot (pynmap/pynmap.py:99-474)."""

    def __init__(self, folder="", snap_name=None, snap_type="ascii", distance=10.0, MLrecipe="SSPs",
                 inclination=0., PA=0., verbose=True, extra_labels=[], ml_function=None,
                 output="numpy"):
        self.verbose = verbose
        self.distance = distance
        self.pc_per_arcsec = self.distance * np.pi / 0.648
        self.MLrecipe = MLrecipe
        self.ml_function = ml_function
        self._PA_orig = PA
        self._inclination_orig = inclination
        self.snap_type = snap_type
        self.extra_labels = extra_labels
        self.output = output
        self._soa = None
        self.folder = folder
        if snap_name is None:
            print('ERROR: no filename for the main input snapshot file provided')
            return
        self.snap_name = snap_name
        in_name = joinpath(self.folder, self.snap_name)
        if not os.path.isfile(in_name):
            print('ERROR: filename {} does not exists'.format(in_name))
            return
        self.read(in_name)
        self._reset_rotations()

    # ------------------------------------------------------------------ construction
    @classmethod
    def from_arrays(cls, pos, vel, mass=None, age=None, MH=None, flux=None, distance=10.0,
                    MLrecipe=None, verbose=False, ml_function=None, output=None, **extra):
        """Build a snapshot from (N, 3) / (3, N)-SoA arrays or tensors instead of a file.
        ``output`` = 'torch' keeps results on the GPU, 'numpy' copies them to the host
        (default: 'torch' if ``pos`` is a CUDA tensor)."""
        self = cls.__new__(cls)
        self.verbose = verbose
        self.distance = distance
        self.pc_per_arcsec = distance * np.pi / 0.648
        self.MLrecipe = MLrecipe
        self.ml_function = ml_function
        self._PA_orig, self._inclination_orig = 0., 0.
        self.snap_type, self.extra_labels = "arrays", list(extra.keys())
        self.folder, self.snap_name = "", "<arrays>"
        self.output = output or ("torch" if engine.is_cuda_tensor(pos) else "numpy")
        for k, v in extra.items():
            setattr(self, k, v)
        self._ingest(pos, vel, mass, age, MH, flux)
        self._reset_rotations()
        return self

    def _ingest(self, pos, vel, mass, age, MH, flux=None):
        dtype = engine.result_dtype(pos, vel)
        self._soa = self._as_soa(pos, dtype), self._as_soa(vel, dtype)
        self.npart = self._soa[0].shape[1]
        dev = self._soa[0].device
        self.mass = engine.to_device(mass, dtype, dev) if mass is not None else torch.ones(self.npart, dtype=dtype, device=dev)
        self.age = engine.to_device(age, dtype, dev) if age is not None else torch.zeros_like(self.mass)
        self.MH = engine.to_device(MH, dtype, dev) if MH is not None else torch.zeros_like(self.mass)
        if flux is not None:
            self.flux = engine.to_device(flux, dtype, dev)
            self.ML = None
        else:
            self.ML = self.compute_ml()
        self.mask = (self.mass == 0)
        self._absmax = {}
        self._acc_pool = {}
        self._cube_plans = {}
        self._epoch = getattr(self, "_epoch", 0) + 1     # what the per-snapshot caches are keyed on (never addresses)
        self.reset()

    @staticmethod
    def _as_soa(a, dtype):
        """-> (3, N) contiguous device tensor; accepts (N, 3) (any memory order) or (3, N)."""
        t = a.detach() if isinstance(a, torch.Tensor) else torch.from_numpy(np.asarray(a))
        if t.dim() != 2 or 3 not in t.shape:
            raise ValueError("positions / velocities must be (N, 3) or (3, N), got %s" % (tuple(t.shape),))
        dev = engine.current_device()
        if t.shape[1] == 3 and t.shape[0] != 3:
            t = t.t()
        elif t.shape == (3, 3):
            t = t.t()   # ambiguous: follow the reference's (N, 3) convention
        out = engine.alloc_soa(3, t.shape[1], dtype, dev)      # rows on 32-byte boundaries whatever N is
        out.copy_(t.to(device=dev))
        return out

    def _check_attrs(self, list_attr=['']):
        return all(hasattr(self, attr) for attr in list_attr)

    def _reset_rotations(self):
        self.PAs = []
        self.inclinations = []

    def read(self, filename):
        """Read the input file (pynmap/pynmap.py:182-232) and upload the particles."""
        if self.verbose:
            print("Opening Input file {0} ".format(filename))
            print("By default, will read all 6 positions/velocities: x, y, z, Vx, Vy, Vz")
        pos, vel, data = dic_io[self.snap_type](filename)
        if pos is None:
            print("ERROR: file may be empty. \n"
                  "       Did not succeed to read positions and velocities.\n"
                  "       - Aborting")
            return
        if data.shape[0] != len(self.extra_labels):
            print("WARNING: cannot convert data beyond positions/velocities/masses")
        else:
            for i, lab in enumerate(self.extra_labels):
                setattr(self, lab, data[i])
        mass = getattr(self, 'mass', None)
        age = getattr(self, 'age', None)
        if age is None and hasattr(self, 'logage'):
            age = 10 ** (self.logage)
        MH = getattr(self, 'MH', None)
        if MH is None and hasattr(self, 'ZH'):
            MH = np.log10(self.ZH) - np.log10(0.19)
        self._ingest(pos, vel, mass, age, MH)

    def compute_ml(self):
        """flux = mass / ML (pynmap/pynmap.py:234-236).  The SSP recipe of the reference
        (misc_functions.compute_ml_ssps: scipy griddata on the E-MILES tables) needs those tables,
        which are data of the reference and are not shipped here: they are looked for in $PNM_SSP_DIR and
        in an installed reference package (ssp.default_ssp_dir), or pass
        ``ml_function=functools.partial(pynmap_b200.ssp.compute_ml_ssps, ssp_dir=<SSPLibraries>)``
        (GPU interpolation), any callable (mass, age, MH) -> M/L, or ``flux=`` to
        from_arrays.  Any other recipe is M/L = 1.  M/L is kept in float64 like the reference's; the
        flux is rounded to the particle dtype."""
        if self.ml_function is not None:
            ml = self.ml_function(self.mass, self.age, self.MH)
            self.ML = engine.to_device(ml, torch.float64, self.mass.device)
        elif self.MLrecipe == "SSPs":
            from . import ssp
            if ssp.default_ssp_dir() is None:
                raise NotImplementedError(
                    "MLrecipe='SSPs' needs the reference's SSP tables: set PNM_SSP_DIR, install the reference "
                    "package, pass ml_function=functools.partial(pynmap_b200.ssp.compute_ml_ssps, ssp_dir=...) or "
                    "precomputed flux; use MLrecipe=None for M/L = 1")
            ml = ssp.compute_ml_ssps(self.mass, self.age, self.MH)       # misc_functions.py:92-113 -> :115-207
            self.ML = engine.to_device(ml, torch.float64, self.mass.device)
        else:
            self.ML = torch.ones_like(self.mass, dtype=torch.float64)
        self.flux = (self.mass.double() / self.ML).to(self.mass.dtype)

    # ------------------------------------------------------------------ rotation state
    def reset(self):
        """Back to the original positions / velocities (pynmap/pynmap.py:238-244): drops the
        recorded matrices; no copy is made because the originals are never overwritten."""
        if getattr(self, "_base", None) is not self._soa:
            self._new_base(self._soa)
        self._chain = []
        self._materialised = None
        self._reset_rotations()

    def rotate(self, PA=0., inclination=90., reset=True, direct=True):
        """Counter-clockwise rotation of PA (around z) then inclination (around x), degrees
        (pynmap/pynmap.py:246-285).  Recorded lazily; see the module docstring."""
        if self._soa is None:
            print('WARNING: no positions - Aborting rotation')
            return
        if reset:
            self.reset()
        LPA = convert_to_list(PA)
        Linclination = convert_to_list(inclination)
        for pa, inclin in zip(LPA, Linclination):
            mat = rotation_matrix(pa, inclin, direct=direct)   # raises for a list of angles, like the reference
            if len(self._chain) >= _lib.MAX_CHAIN:
                self._rebase()
            self._chain.append(mat)
            self._materialised = None
            self.PAs.append(pa)
            self.inclinations.append(inclin)
        if self.verbose:
            print('Rotations done!')

    def _materialise(self):
        if self._materialised is None:
            pos, vel = self._base
            for mat in self._chain:
                pos = engine.rotate_soa((pos[0], pos[1], pos[2]), mat)
                vel = engine.rotate_soa((vel[0], vel[1], vel[2]), mat)
            self._materialised = (pos, vel)
        return self._materialised

    def _rebase(self):
        self._new_base(self._materialise())
        self._chain = []

    def _new_base(self, base):
        """The arrays the fused kernels read have changed: everything derived from them (velocity bounds, the
        memory-order estimate, cube plans, fixed-point scales) is forgotten."""
        self._base = base
        self._epoch = getattr(self, "_epoch", 0) + 1
        self._absmax = {}
        self._cube_plans = {}

    def _out(self, t):
        return t if self.output == "torch" else t.cpu().numpy()

    @property
    def pos(self):
        """(N, 3) rotated positions, column-contiguous like the reference's (rotation.py:115)."""
        if self._soa is None:
            return None
        return self._out(self._materialise()[0].t())

    @property
    def vel(self):
        if self._soa is None:
            return None
        return self._out(self._materialise()[1].t())

    def info(self):
        print("Input file name: {0} with type {1}".format(self.snap_name, self.snap_type))
        print("Number of particles: {0}".format(self.npart))
        print("Original PA={0} and inclination={1}".format(self._PA_orig, self._inclination_orig))
        print("Applied rotations: PAs        Inclinations")
        if len(self.PAs) > 0:
            for k, (pa, inclin) in enumerate(zip(self.PAs, self.inclinations)):
                print("{0}          {1:8.2f}       {2:8.2f}".format(k + 1, pa, inclin))
        else:
            print("No rotations applied (yet)")

    def _select_weight(self, weight=None):
        if weight is None:
            return weight
        elif weight == "mass":
            return self.mass
        elif weight == "lum":
            return self.flux
        print("ERROR: please specify a weight function as 'mass' or 'lum'")
        return None

    # ------------------------------------------------------------------ the fused path
    def _weights_on_device(self, weights):
        if weights is None:
            return None
        dtype = self._soa[0].dtype
        w = engine.to_device(weights, dtype, self._soa[0].device).reshape(-1)
        if w.numel() != self.npart:
            raise ValueError("weights have %d entries for %d particles" % (w.numel(), self.npart))
        return w

    def _max_of(self, key, tensor):
        """max |tensor|, cached under ``key`` -- only snapshot-owned arrays have a key (the caches are cleared whenever
        those arrays change: _ingest, _new_base); anything the caller passed in is measured on every call."""
        if key is None:
            return engine.absmax(tensor)
        if key not in self._absmax:
            self._absmax[key] = engine.absmax(tensor)
        return self._absmax[key]

    def _weights_key(self, w):
        """Cache key of a weights tensor: the attribute name when it IS the snapshot's mass / flux, else None."""
        if w is None:
            return ("unit", 0)
        for name in ("mass", "flux"):
            t = getattr(self, name, None)
            if isinstance(t, torch.Tensor) and t.data_ptr() == w.data_ptr() and t.numel() == w.numel() and t.dtype == w.dtype:
                return (name, t._version)          # in-place torch updates of the attribute bump _version
        return None

    def _bounds(self, w):
        """(max |w|, bound on |v_los|) for the fixed-point scales."""
        vel = self._base[1]
        max_v = 2.0 * max(self._max_of(("v", k), vel[k]) for k in range(3))  # |row . v| <= sqrt(3) max|v_k|
        wkey = self._weights_key(w)
        max_w = 1.0 if w is None else self._max_of(None if wkey is None else ("w", wkey), w)
        return max_w, max_v

    def _fixed_scales(self, acc, w, n_total=None, reduce_max=None, mean_abs=None):
        if acc.code != _lib.ACC_FIXED:
            return None
        max_w, max_v = self._bounds(w)
        if reduce_max is not None:            # sharded runs agree on global bounds
            max_w, max_v = reduce_max(max_w, max_v)
        if mean_abs is not None:              # runs through k_project_cube: typical terms near 2^24 (engine.cube_fixed_scales)
            if reduce_max is not None:
                mean_abs = reduce_max(*mean_abs)
            return engine.cube_fixed_scales(n_total or self.npart, max_w, max_v, mean_abs)
        return engine.fixed_scales(n_total or self.npart, max_w, max_v)

    def _run_fused(self, grid, w, sums, acc, scales, mats=None, nchain=None, nviews=1, mask=None, mask_mode=0,
                   recycled=None, plan=None, slab_multiple=1):
        pos, vel = self._base
        if mats is None:
            mats = self._chain if self._chain else [np.eye(3, dtype=np.float32)]
            nchain = len(mats)
        maps, cube = recycled if recycled is not None else engine.new_accumulators(grid, sums, acc, nviews, slab_multiple)
        soa = (pos[0], pos[1], pos[2], vel[0], vel[1], vel[2])
        if plan is not None:
            engine.project_cube(grid, plan, soa, w, mats[0], acc, scales, cube)
        else:
            engine.project_bin(grid, soa, w, np.stack(mats), nchain, nviews, sums, acc, scales, maps, cube, mask, mask_mode)
        return maps, cube

    def _cube_plan(self, grid, w):
        """The plan of the privatised scatter (engine.CubePlan) for the current view, or None when that kernel does
        not apply (float64 particles, compounded rotations, unaligned arrays).  One plan per (grid, view), built on
        the device from a sample of the particles; forgotten when the arrays change."""
        if len(self._chain) > 1:
            return None
        pos, vel = self._base
        soa = (pos[0], pos[1], pos[2], vel[0], vel[1], vel[2])
        if not engine.cube_path_ok(grid, soa, w):
            return None
        mat = self._chain[0] if self._chain else np.eye(3, dtype=np.float32)
        # One entry per (grid, view, weights attribute).  The plan also carries the typical term sizes that set the
        # shared-memory / fixed-point scales, so it must follow the CONTENTS of the weights: the snapshot's own mass /
        # flux are tracked through their torch version counter; weights the caller passes in cannot be tracked and
        # get their plan re-built (0.4 ms, same allocation) on every call -- nothing is keyed on addresses.
        wkey = self._weights_key(w)
        key = (id(grid), np.asarray(mat, dtype=np.float32).tobytes(), wkey[0] if wkey else "caller")
        entry = self._cube_plans.get(key)
        if entry is None:
            if len(self._cube_plans) >= 16:
                self._cube_plans.clear()
            plan = engine.CubePlan(grid).build(soa, w, mat)
            entry = self._cube_plans[key] = (plan, grid, wkey)  # the grid is held so that its id stays unique
        elif wkey is None or entry[2] != wkey:
            entry[0].build(soa, w, mat)
            entry = self._cube_plans[key] = (entry[0], grid, wkey)
        return entry[0]

    def _mesh(self, grid, share=False):
        return grid.mesh(self.output == "torch", share)

    @staticmethod
    def _mask_arg(kwargs, mode):
        mask = kwargs.pop("mask", None)
        if mask is None:
            return None, _lib.MASK_NONE
        return engine.to_device(mask, torch.uint8).reshape(-1), mode

    def velocity_moments(self, data=None, weights=None, weight_name=None, limXY=default_limXY,
                         nXY=default_npoint, **kwargs):
        """Weighted mass, mean velocity and dispersion maps (pynmap/pynmap.py:319-340)."""
        if self._soa is None:
            print("WARNING: no positions - Aborting")
            return SnapMap(name="EmptyMap")
        if weight_name is not None:
            weights = self._select_weight(weight_name)
        if data is None:
            print("WARNING: using z component of V as data")
            n2 = engine.expand_nxy(nXY)
            if n2 is None:
                print("ERROR: dimension of n should be 1 or 2")
                X = Y = M = V = S = 0
            else:
                acc = engine.AccMode(kwargs.pop("accumulate", "f64"))
                reducer = kwargs.pop("reducer", None)        # sharded.GridReducer: partial grids summed over the ranks
                mask, mode = self._mask_arg(kwargs, _lib.MASK_AND_NONFINITE)
                grid = engine.get_grid(limXY, n2[0], n2[1])
                w = self._weights_on_device(weights)
                sums = _lib.SUM_W | _lib.SUM_WD | _lib.SUM_WD2
                if reducer is not None and acc.code == _lib.ACC_FIXED:   # every rank must quantise identically
                    scales = self._fixed_scales(acc, w, reducer.n_total(self.npart), reducer.reduce_max)
                else:
                    scales = self._fixed_scales(acc, w)
                maps, _ = self._run_fused(grid, w, sums, acc, scales, mask=mask, mask_mode=mode)
                if reducer is not None:
                    engine.fold_maps(grid, maps, acc, clear=False)
                    if not reducer(*engine.reduce_span(grid, maps, None)):
                        return None                           # another rank holds the total
                    sums |= _lib.MAPS_FOLDED
                M, V, S = (self._out(t) for t in engine.finalize_vmaps(grid, maps, None, sums, acc, scales))
                X, Y = self._mesh(grid)
        else:
            pos = self._materialise()[0]
            X, Y, M, V, S = points_2vmaps(pos[0], pos[1], data, weights=weights, limXY=limXY, nXY=nXY, **kwargs)
            X, Y, M, V, S = (self._conv(a) for a in (X, Y, M, V, S))
        mydict = {"X": X, "Y": Y, "V": V, "M": M, "S": S}
        mydict["PAs"] = self.PAs
        mydict["inclinations"] = self.inclinations
        return SnapMap(name="Velocity moments", input_dic=mydict)

    # (``velocity_moments(..., reducer=sharded.GridReducer())``: every rank bins its own particles, the partial maps are
    #  summed with one reduce / all-reduce -- BASELINE.json config 3; ranks that do not hold the total get None)

    def amr_velocity_moments(self, data=None, weights=None, weight_name=None, box=150.,
                             limXY=default_limXY, nXY=default_npoint, sigfac=10, smear=False, spread=True,
                             **kwargs):
        """Velocity moments for AMR gas (pynmap/pynmap.py:342-419): every cell is replicated 10 times
        inside its footprint box / 2**level (``spread``; numpy's global generator supplies the
        deviates exactly as in the reference, see amr.py) and gridded with points_2vmaps.

        ``smear=True`` (per-level maps convolved with astropy's ``convolve_fft`` and an explicitly
        sized, possibly even ``Gaussian2DKernel``, pynmap.py:379-409) is not available: astropy is
        not part of this stack and that branch cannot be pinned to it."""
        if not self._check_attrs(['level']):
            print("ERROR: missing some input info in class")
            return None
        if smear:
            raise NotImplementedError("amr_velocity_moments(smear=True) needs astropy.convolution.convolve_fft")
        if weight_name is not None:
            weights = self._select_weight(weight_name)
        pos, vel = self._materialise()
        if data is None:
            data = vel[2]
        if weights is None:
            weights = torch.ones(self.npart, dtype=pos[0].dtype, device=pos[0].device)     # np.ones_like(self.mass)
        if spread:
            from .amr import spread_amr
            [x, y], [weightr], [datar, _] = spread_amr([pos[0], pos[1]], self.level, list_val=[weights],
                                                       list_const=[data, self.level], nsample=10, box=box,
                                                       stretch=[1., 1.])
        else:
            x, y, datar, weightr = pos[0], pos[1], data, weights
        X, Y, M, V, S = points_2vmaps(x, y, datar, weights=weightr, limXY=limXY, nXY=nXY)
        X, Y, M, V, S = (self._conv(a) for a in (X, Y, M, V, S))
        mydict = {"X": X, "Y": Y, "V": V, "M": M, "S": S}
        mydict["PAs"] = self.PAs
        mydict["inclinations"] = self.inclinations
        return SnapMap(name="Velocity moments", input_dic=mydict)

    def get_tensor_profiles(self, nbins=100, dz=np.array([-0.2, 0.2]), Rmin=1.e-2, Rmax=None, Vmax=None,
                            plot=False, saveplot=False, **kwargs):
        """Radial profiles of the three velocity means and dispersions in cylindrical coordinates
        (pynmap/pynmap.py:457-474 -> grid.point_2vprofiles)."""
        pos, vel = self._materialise()
        Rbin, v, s = point_2vprofiles(pos[0], pos[1], pos[2], vel[0], vel[1], vel[2], nbins=nbins, dz=dz, Rmin=Rmin,
                                      Rmax=Rmax, Vmax=Vmax, plot=plot, saveplot=saveplot, **kwargs)
        return Rbin, self._conv(v), self._conv(s)

    def _conv(self, a):
        if isinstance(a, torch.Tensor):
            return self._out(a)
        return a

    def create_map(self, data=None, weights=None, weight_name=None, limXY=default_limXY,
                   nXY=default_npoint, **kwargs):
        """Weighted mean map of an arbitrary per-particle quantity (pynmap/pynmap.py:421-437)."""
        if weight_name is not None:
            weights = self._select_weight(weight_name)
        if data is None:
            data = self.mass
        pos = self._materialise()[0]
        X, Y, W, D, DW = points_2map(pos[0], pos[1], data, weights=weights, limXY=limXY, nXY=nXY, **kwargs)
        X, Y, W, D, DW = (self._conv(a) for a in (X, Y, W, D, DW))
        mydict = {"X": X, "Y": Y, "W": W, "D": D, "DW": DW}
        mydict["PAs"] = self.PAs
        mydict["inclinations"] = self.inclinations
        return SnapMap(name="map", input_dic=mydict)

    def create_losvds(self, data=None, weights=None, weight_name=None, limXY=default_limXY,
                      nXY=default_npoint, limV=default_limV, nV=default_nVpoint, **kwargs):
        """Line-of-sight velocity distributions on a grid (pynmap/pynmap.py:439-455)."""
        if weight_name is not None:
            weights = self._select_weight(weight_name)
        if data is None:
            n2 = engine.expand_nxy(nXY)
            if n2 is None:
                print("ERROR: dimension of n should be 1 or 2")
                return 0, 0, 0, 0, 0
            acc = engine.AccMode(kwargs.pop("accumulate", "f64"))
            mask, mode = self._mask_arg(kwargs, _lib.MASK_PLAIN)
            w = self._weights_on_device(weights)
            if mask is not None and w is not None and bool(mask.any()):
                raise ValueError("The weights and list don't have the same length.")   # grid.py:251-259
            grid = engine.get_grid(limXY, n2[0], n2[1], limV, int(nV))
            scales = self._fixed_scales(acc, w)
            _, cube = self._run_fused(grid, w, _lib.SUM_CUBE, acc, scales, mask=mask, mask_mode=mode)
            px, py, pv = grid.centres()
            mylosvd = LOSVD(px, py, pv, self._out(engine.finalize_cube(grid, cube, acc, scales)))
        else:
            pos = self._materialise()[0]
            mylosvd = points_2losvds(pos[0], pos[1], data, weights=weights, limXY=limXY, nXY=nXY,
                                     limV=limV, nV=nV, **kwargs)
            if isinstance(mylosvd, LOSVD):
                mylosvd.losvd = self._conv(mylosvd.losvd)
        mylosvd.PAs = self.PAs
        mylosvd.inclinations = self.inclinations
        return mylosvd

    # ------------------------------------------------------------------ additions (single read)
    def project(self, weights=None, weight_name=None, limXY=default_limXY, nXY=default_npoint,
                limV=default_limV, nV=default_nVpoint, accumulate="f64", reducer=None, post_stream=None,
                privatise=None, deferred=False):
        """velocity_moments + create_losvds from ONE pass over the particles: returns
        (SnapMap, LOSVD) identical to calling the two methods separately.  The map's sum of
        weights is recovered from the cube (plus an out-of-range remainder), which saves one
        atomic per particle.  ``reducer(maps, cube)`` is the multi-GPU hook (sharded.py).

        ``privatise``: None (default) lets float32 snapshots in no particular memory order run through the
        shared-memory privatised kernel (csrc/pnm_cube.cuh; accumulate='f64' then rounds the privatised terms to
        2^-25 of the typical term, accumulate='fixed' uses coarser scales -- see engine.cube_fixed_scales); False
        forces the plain reduction kernel.

        ``deferred=True``: only the scatter kernel (and, with a slab reducer, the start of the exchange between ranks)
        is launched; returns a PendingProjection whose ``result()`` launches the rest and returns (SnapMap, LOSVD).
        Calling ``result()`` after the NEXT project(deferred=True) lets the exchange of this step overlap the next
        scatter kernel (sharded.PeerSlabReducer).

        ``post_stream``: a torch.cuda.Stream on which everything AFTER the scatter kernel runs
        (replica fold, the reduce over ranks, finalisation).  The scatter of the next call then
        overlaps this call's collective and post-processing; the returned tensors belong to
        ``post_stream`` -- keep working on them inside ``with torch.cuda.stream(post_stream)``
        and synchronise it before reading them elsewhere."""
        if weight_name is not None:
            weights = self._select_weight(weight_name)
        n2 = engine.expand_nxy(nXY)
        acc = engine.AccMode(accumulate)
        grid = engine.get_grid(limXY, n2[0], n2[1], limV, int(nV))
        w = self._weights_on_device(weights)
        sums = _lib.SUM_W | _lib.SUM_WD | _lib.SUM_WD2 | _lib.SUM_CUBE | _lib.W_FROM_CUBE
        slabs = bool(getattr(reducer, "scatter_slabs", False))      # sharded.SlabReducer: reduce-scatter along the record slabs
        slab_multiple = reducer.world if slabs else 1
        if slabs or not self._memory_ordered(grid, reducer):
            # particles in no particular order: all three sums of a particle into ONE sector of the interleaved cube.
            # Spatially ordered snapshots (tree / space-filling-curve order) keep the plain cube + replicated maps,
            # where consecutive particles of a pixel merge in registers and hot pixels are spread over replicas.
            sums |= _lib.CUBE_MOMENTS
        want_plan = (sums & _lib.CUBE_MOMENTS) and acc.privatise and privatise is not False
        plan = self._cube_plan(grid, w) if want_plan else None
        if plan is not None:      # sharded runs: the collective of the previous step runs next to this scatter
            plan.reserve_sms(getattr(reducer, "reserve_sms", 0) if post_stream is not None else 0)
        scales = None
        if acc.code == _lib.ACC_FIXED:
            # sharded fixed-point runs agree on the global particle count and bounds.  Cached for the snapshot's own
            # weights only (the two tiny all-reduces end in a host read and would serialise every step); weights the
            # caller passes in are measured again on every call -- their contents may have changed.
            wkey = self._weights_key(w)
            key = ("fixed", wkey, id(plan), id(reducer))
            if wkey is None or key not in self._absmax:
                n_total, reduce_max = (reducer.n_total(self.npart), reducer.reduce_max) if reducer else (None, None)
                mean_abs = plan.info()["mean_abs"] if plan is not None else None
                val = (self._fixed_scales(acc, w, n_total, reduce_max, mean_abs), reducer)
                if wkey is None:
                    scales = val[0]
                else:
                    self._absmax[key] = val
            if wkey is not None:
                scales = self._absmax[key][0]
            if plan is not None and not engine.cube_scales_ok(scales):
                plan = None
        self.last_fixed_scales = scales        # (what a checker needs to reproduce the int64 sums)
        if post_stream is not None and (deferred or hasattr(reducer, "acquire")):
            raise ValueError("post_stream cannot be combined with deferred=True or a reducer that owns its accumulators "
                             "(sharded.PeerSlabReducer): use project(deferred=True)")
        if post_stream is None:
            recycled = (None, reducer.acquire(grid, acc)) if slabs and hasattr(reducer, "acquire") else None
            maps, cube = self._run_fused(grid, w, sums, acc, scales, plan=plan, slab_multiple=slab_multiple, recycled=recycled)
            token = reducer.begin(cube, grid, acc) if slabs else None
            pending = PendingProjection(self, grid, maps, cube, sums, acc, scales, reducer, token)
            return pending if deferred else pending.result()
        # pipelined: the accumulators are recycled -- cleared on the post-processing stream once their results are
        # out, so that no memset sits between two scatter kernels on the caller's stream
        pool = self._acc_pool.setdefault((id(grid), sums, acc.code, slab_multiple), [])
        recycled = None
        if len(pool) >= 2:                                   # the older of two sets: its clearing finished a step ago
            maps, cube, cleared = pool.pop(0)
            torch.cuda.current_stream().wait_event(cleared)
            recycled = (maps, cube)
        maps, cube = self._run_fused(grid, w, sums, acc, scales, recycled=recycled, plan=plan, slab_multiple=slab_multiple)
        scattered = torch.cuda.Event()
        scattered.record()                                   # on the caller's (scatter) stream
        post_stream.wait_event(scattered)
        for t in (maps, cube):                               # keep the storage alive for the other stream
            if t is not None:
                t.record_stream(post_stream)
        with torch.cuda.stream(post_stream):
            out = self._project_post(grid, maps, cube, sums, acc, scales, reducer)
            cube.zero_()
            if maps is not None:
                maps.zero_()
            cleared = torch.cuda.Event()
            cleared.record()
        pool.append((maps, cube, cleared))
        return out

    def _memory_ordered(self, grid, reducer=None):
        """True when neighbours in memory are neighbours in space (median distance between consecutive particles
        below one pixel), estimated once per snapshot and pixel size on a strided sample of 65536 pairs by one small
        kernel of the library (pnm_order_probe: a count of close pairs, no sort).  Sharded
        runs agree on one answer (any rank ordered -> all use the plain layout): the partial grids are summed."""
        pos = self._base[0]
        pix = min((grid.limXY[1] - grid.limXY[0]) / max(grid.nx - 1, 1), (grid.limXY[3] - grid.limXY[2]) / max(grid.ny - 1, 1))
        key = ("ordered", float(pix), id(reducer))
        if key not in self._absmax:
            ordered = False
            if self.npart >= 8192:
                close, looked = engine.order_probe(pos[0], pos[1], pos[2], pix)
                ordered = looked > 0 and 2 * close > looked      # median distance below one pixel, without a sort
            if reducer is not None:
                ordered = bool(reducer.reduce_max(float(ordered))[0] > 0.5)
            self._absmax[key] = (ordered, reducer)       # (the reducer is held so that its id stays unique)
        return self._absmax[key][0]

    def _project_post(self, grid, maps, cube, sums, acc, scales, reducer):
        """Everything after the scatter of project(): fold + reduce over ranks, finalisation."""
        if getattr(reducer, "scatter_slabs", False):
            return self._project_post_slabs(grid, cube, acc, scales, reducer, reducer.begin(cube, grid, acc))
        if reducer is not None:
            if not sums & _lib.CUBE_MOMENTS:                   # (the interleaved layout has a single map replica)
                engine.fold_maps(grid, maps, acc, clear=False) # single-use accumulators: nothing adds to them again
            keep = reducer(*engine.reduce_span(grid, maps, cube))
            if not keep:
                return None, None
            sums |= _lib.MAPS_FOLDED
        M, V, S = (self._out(t) for t in engine.finalize_vmaps(grid, maps, cube, sums, acc, scales))
        X, Y = self._mesh(grid)
        newmap = SnapMap(name="Velocity moments", input_dic={"X": X, "Y": Y, "V": V, "M": M, "S": S,
                                                              "PAs": self.PAs, "inclinations": self.inclinations})
        px, py, pv = grid.centres()
        mylosvd = LOSVD(px, py, pv, self._out(engine.finalize_cube(grid, cube, acc, scales, sums)))
        mylosvd.PAs, mylosvd.inclinations = self.PAs, self.inclinations
        return newmap, mylosvd

    def _project_post_slabs(self, grid, cube, acc, scales, reducer, token):
        """Sharded post-processing (sharded.SlabReducer): reduce-scatter of the record slabs, every rank finalises
        its own velocity bins; the maps come from an all-reduce of the per-rank partial sums (3 values per pixel).
        -> (SnapMap, LOSVD of this rank's velocity bins) on every rank."""
        g0, g1 = reducer.slab_range(grid)
        mine = reducer.finish(token)
        sums3 = reducer.map_sums(token, mine, grid, acc)
        M, V, S = (self._out(t) for t in engine.finalize_sums(grid, sums3, acc, scales))
        X, Y = self._mesh(grid, share=True)
        newmap = SnapMap(name="Velocity moments", input_dic={"X": X, "Y": Y, "V": V, "M": M, "S": S,
                                                              "PAs": self.PAs, "inclinations": self.inclinations})
        px, py, pv = grid.centres()
        iv0, iv1 = engine.slab_vrange(grid, g0, g1)
        mylosvd = LOSVD(px, py, pv[iv0:iv1], self._out(engine.finalize_cube_slabs(grid, mine, g0, g1, acc, scales)))
        mylosvd.PAs, mylosvd.inclinations = self.PAs, self.inclinations
        mylosvd.vrange = (iv0, iv1)                            # which velocity bins of the whole cube these are
        return newmap, mylosvd

    def project_views(self, PAs, inclinations, weights=None, weight_name=None, limXY=default_limXY,
                      nXY=default_npoint, accumulate="f64", direct=True, views_per_launch=1):
        """Moment maps for many viewing angles (the reference needs a Python loop of rotate() +
        velocity_moments(), pynmap.py:277-282).  Each view starts from the original configuration,
        like rotate(reset=True).  Returns a list of SnapMap.

        ``views_per_launch`` > 1 bins several views per read of the particles (pnm_project_bin with
        nviews > 1).  The default is one launch per view: re-reading the particles costs far less than
        the scatter itself, and the multi-view kernel gives up the one-request-per-particle packing of
        the single-view maps path -- measured 14.9 ms (64 launches) vs 24.5 / 25.2 / 43.2 ms (2 / 8 / 64
        views per launch) for 2e7 particles x 64 views (profiles/r02_micro_views_bench.log)."""
        if weight_name is not None:
            weights = self._select_weight(weight_name)
        n2 = engine.expand_nxy(nXY)
        acc = engine.AccMode(accumulate)
        grid = engine.get_grid(limXY, n2[0], n2[1])
        w = self._weights_on_device(weights)
        sums = _lib.SUM_W | _lib.SUM_WD | _lib.SUM_WD2
        saved = (self._base, self._chain)
        self._base, self._chain = self._soa, []
        try:
            scales = self._fixed_scales(acc, w)
            out = []
            pas = [np.float32(p) for p in PAs]
            incs = [np.float32(i) for i in inclinations]
            per = max(1, min(int(views_per_launch), _lib.MAX_VIEWS))
            for s in range(0, len(pas), per):
                mats = [rotation_matrix(p, i, direct=direct) for p, i in zip(pas[s:s + per], incs[s:s + per])]
                maps, _ = self._run_fused(grid, w, sums, acc, scales, mats=mats, nchain=1, nviews=len(mats))
                X, Y = self._mesh(grid)
                for k in range(len(mats)):
                    view = maps[k * grid.maps_elems:(k + 1) * grid.maps_elems]
                    M, V, S = (self._out(t) for t in engine.finalize_vmaps(grid, view, None, sums, acc, scales))
                    out.append(SnapMap(name="Velocity moments", input_dic={
                        "X": X, "Y": Y, "V": V, "M": M, "S": S, "PAs": [pas[s + k]], "inclinations": [incs[s + k]]}))
        finally:
            self._base, self._chain = saved
        return out
