/*
 * pnm_b200.h -- C ABI of libpnm_b200.so: pynmap's particle-projection hot path on B200 (sm_100a).
 *
 * The reference (emsellem/pynmap) has no FFI layer: the path is four Python functions whose
 * arithmetic lives in numpy / astropy library calls.  Each entry point below replaces the BODY
 * of one of those calls; the Python host layer (the pynmap_b200 package) keeps the reference's
 * signatures and binds these symbols with ctypes.  Citations are file:line in emsellem/pynmap.
 *
 *   pnm_rotate            np.dot(mat, pos.T).T                 pynmap/rotation.py:115-117, :67-69
 *   pnm_grid_create       np.linspace bin edges                pynmap/grid.py:121-124, :196-199, :243-245
 *   pnm_project_bin       rotate_mat + 3x np.histogram2d       pynmap/rotation.py:73-119 + grid.py:126-131
 *                         + np.histogramdd                     + grid.py:252-259 (one read of the particles)
 *   pnm_bin_points        np.histogram2d x3 / x2, histogramdd  pynmap/grid.py:126-131, :200-201, :259
 *   pnm_finalize_vmaps    V = S1/S0, S = sqrt(S2/S0 - V*V)     pynmap/grid.py:133-138
 *   pnm_finalize_map      D = DW/W                             pynmap/grid.py:203-205
 *   pnm_finalize_cube     weights-histogram -> float64 cube    pynmap/grid.py:259
 *   pnm_smooth_cube       astropy convolve per velocity slice  pynmap/losvd.py:127-135
 *   pnm_cube_moments      moment012 per spaxel                 pynmap/losvd.py:41-51, misc_functions.py:70-90
 *   pnm_radial_profiles   point_2vprofiles                     pynmap/grid.py:16-68
 *   pnm_amr_jitter/repeat spread_amr                           pynmap/misc_io.py:184-224
 *   pnm_fit_gh            fitgh (mpfit) per spaxel             pynmap/losvd.py:53-67, fit_losvd.py:57-235
 *   pnm_absmax            (new) bound for the fixed-point scale
 *   pnm_project_cube      pnm_project_bin for the hot configuration (float32 particles, one view, maps + cube),
 *   pnm_cube_plan_*       hot pixels / voxels accumulated in shared memory first (same sums, same layout)
 *   pnm_project_host      snapshot.rotate + velocity_moments + create_losvds from HOST arrays
 *                                                              pynmap/pynmap.py:246-285, :319-340, :439-455
 *
 * Conventions
 *   - every function returns 0 on success, a negative PNM_E* code otherwise; the message is
 *     available from pnm_last_error() (thread local);
 *   - all particle / accumulator / output pointers are DEVICE pointers unless the name ends in
 *     _host; the library never allocates or frees caller-visible memory.  The only internal
 *     allocations are the small edge tables owned by a pnm_grid (and the staging ring of
 *     pnm_project_host);
 *   - all work is enqueued on the cudaStream_t passed as `stream` (a void* here so that the
 *     header needs no CUDA include); nothing synchronises except pnm_project_host;
 *   - there is no CPU fallback: every entry point fails with PNM_ECUDA when no sm_100 device
 *     is usable.
 */
#ifndef PNM_B200_H
#define PNM_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PNM_VERSION 100 /* 0.1.0 */

/* error codes */
#define PNM_OK 0
#define PNM_EINVAL (-1)   /* bad argument (null pointer, size, flag combination) */
#define PNM_EDTYPE (-2)   /* unsupported dtype */
#define PNM_ECUDA (-3)    /* CUDA runtime error (message has the cudaError string) */
#define PNM_ELIMIT (-4)   /* size above a compiled-in limit (edges, taps, chain length) */

/* particle array dtypes (np.result_type(float32, input) of the reference, rotation.py:115) */
#define PNM_F32 0
#define PNM_F64 1

/* accumulator arithmetic */
#define PNM_ACC_F64 0   /* float64 atomics: <=1e-12 relative vs the sequential float64 sums    */
#define PNM_ACC_FIXED 1 /* int64 fixed point: order independent, bit-identical on any #GPUs   */

/* which sums one call accumulates */
#define PNM_SUM_W 1     /* sum w           -> maps slot 0                                     */
#define PNM_SUM_WD 2    /* sum w*d         -> maps slot 1   (d = v_los or user data)          */
#define PNM_SUM_WD2 4   /* sum w*(d*d)     -> maps slot 2                                     */
#define PNM_SUM_CUBE 8  /* (x, y, d) weighted histogram -> cube                               */
/* With PNM_SUM_W|PNM_SUM_CUBE and PNM_W_FROM_CUBE the map's sum w is not accumulated
 * directly: particles inside the V range go to the cube only, the others to maps slot 0,
 * and pnm_finalize_vmaps adds the cube's V-sum back.  One atomic fewer per particle.       */
#define PNM_W_FROM_CUBE 16
/* With PNM_W_FROM_CUBE: the cube accumulator uses the INTERLEAVED layout [ceil(nv/2)][ix][iy][4] =
 * { w of velocity bin 2g, w of bin 2g+1, sum w*d, sum w*d*d } (pnm_cube_moments_acc_elems elements), so that all
 * three sums of a particle inside the V range fall into one 32-byte sector -- one L2 request per particle.  The maps
 * accumulator then only receives the particles outside the V range and is a single replica (ny*nx*PNM_MAP_SLOTS
 * elements, no pnm_fold_maps needed); pnm_finalize_vmaps / pnm_finalize_cube take the same flag.   */
#define PNM_CUBE_MOMENTS 256
/* Mixed input dtypes (pnm_bin_points with dtype = PNM_F64 only).  numpy evaluates `wm * vm` and
 * `wm * vm**2` (grid.py:128-131, :201) in the result type of the two operands, which is float32 when
 * both are float32 even if the coordinates are float64.  Arrays are still passed as float64:
 *   PNM_VAL_D32: d holds float32 values, d*d is rounded in float32;
 *   PNM_VAL_W32: w holds float32 values too; with both set w*d and w*(d*d) are rounded in float32.  */
#define PNM_VAL_D32 64
#define PNM_VAL_W32 128
/* pnm_finalize_vmaps only: the first map replica already holds the totals (pnm_fold_maps, possibly
 * followed by a reduction over ranks); the other replicas are not read.                          */
#define PNM_MAPS_FOLDED 32

/* mask semantics of the reference */
#define PNM_MASK_NONE 0
#define PNM_MASK_AND_NONFINITE 1 /* drop iff mask && !isfinite(d)   grid.py:113-119, :188-194 */
#define PNM_MASK_PLAIN 2         /* drop iff mask                   grid.py:251-254           */

#define PNM_MAX_EDGES 4097 /* bins + 1 per axis                                              */
#define PNM_MAX_CHAIN 8    /* compounded rotations applied in registers (rotate(reset=False)) */
#define PNM_MAX_VIEWS 64   /* viewing angles per particle read                               */
#define PNM_MAX_TAPS 255   /* Gaussian taps per axis                                         */
#define PNM_MAP_SLOTS 4    /* interleaved accumulator slots per pixel: S0, S1, S2, reserved  */

typedef struct pnm_grid pnm_grid; /* opaque: edge tables on one device */

int pnm_version(void);
const char* pnm_last_error(void);
/* number of usable sm_100 devices (0 if none / no driver) */
int pnm_device_count(void);

/* ---- a1: materialising rotation (rotation.py:115-117) -------------------------------------
 * out = fl(fma(m[3r+2], z, fma(m[3r+1], y, fl(m[3r]*x)))) per row r, in `dtype` arithmetic,
 * mat9_host = row-major float32 3x3 on the HOST.  In-place (ox==x ...) is allowed.          */
int pnm_rotate(const void* x, const void* y, const void* z, void* ox, void* oy, void* oz,
               int64_t n, int dtype, const float* mat9_host, void* stream);

/* ---- grid (edge tables) -------------------------------------------------------------------
 * edges_*_host: float64 arrays of n*+1 monotonically increasing edges built by the host with
 * the reference's own numpy expression (grid.py:121-124).  nv == 0 / edges_v_host == NULL: no
 * velocity axis.  Bin rule = numpy histogramdd: edges[i] <= x < edges[i+1], last edge
 * inclusive, NaN / out-of-range dropped; comparisons exact in float64.                       */
int pnm_grid_create(int32_t nx, const double* edges_x_host, int32_t ny, const double* edges_y_host,
                    int32_t nv, const double* edges_v_host, pnm_grid** out);
int pnm_grid_destroy(pnm_grid* g);

/* Accumulators are OPAQUE 8-byte-element buffers owned by the caller (zero them before the first
 * call; calls add).  Their internal layout is chosen for the L2 reduction units, not for the
 * reader: the maps accumulator holds pnm_grid_map_replicas() copies of [iy][ix][4] that are summed
 * at finalisation, the cube accumulator is velocity-major.  Sizes per view, in 8-byte elements:  */
int64_t pnm_maps_acc_elems(const pnm_grid* g); /* replicas*ny*nx*PNM_MAP_SLOTS */
int64_t pnm_cube_acc_elems(const pnm_grid* g); /* nx*ny*nv                     */
int64_t pnm_cube_moments_acc_elems(const pnm_grid* g); /* (ceil(nv/2)+1)*nx*ny*4 with PNM_CUBE_MOMENTS */
int pnm_grid_map_replicas(const pnm_grid* g);
/* Sums the map replicas into the first one, so that a multi-GPU reduction only has to move the
 * first ny*nx*PNM_MAP_SLOTS elements.  clear_others != 0 also zeroes replicas 1.. (the buffer can
 * then keep accumulating); with 0 they are left as they are and the buffer must be finalised with
 * PNM_MAPS_FOLDED in `sums` (half the memory traffic: the fold runs next to the following step's
 * scatter in the sharded pipeline).                                                              */
int pnm_fold_maps(const pnm_grid* g, void* acc_maps, int acc_mode, int clear_others, void* stream);

/* ---- fused rotate + project + bin (a1+a2+a3+a4) ---------------------------------------------
 * For each of `nviews` views: apply `nchain` float32 matrices in sequence to (x,y,z) and
 * (vx,vy,vz) (each step rounded like pnm_rotate), take x' = pos'[0], y' = pos'[1],
 * d = vel'[2], then accumulate the sums selected by `sums` with weights w (NULL = 1):
 *   acc_maps [view][pnm_maps_acc_elems]   (double or int64 by acc_mode; [view][ny*nx*PNM_MAP_SLOTS] with PNM_CUBE_MOMENTS)
 *   acc_cube [view][pnm_cube_acc_elems]   ([view][pnm_cube_moments_acc_elems] with PNM_CUBE_MOMENTS)
 * mats_host: float32 [nviews][nchain][9] on the HOST.  scales3_host (FIXED only): the three
 * power-of-two scales for w, w*d, w*d*d.  Accumulators must be zeroed by the caller (or hold
 * earlier partial sums: calls add).  mask: optional uint8 per particle.                       */
int pnm_project_bin(const pnm_grid* g, int64_t n, int dtype,
                    const void* x, const void* y, const void* z,
                    const void* vx, const void* vy, const void* vz,
                    const void* w, const uint8_t* mask, int mask_mode,
                    const float* mats_host, int nchain, int nviews,
                    int sums, int acc_mode, const double* scales3_host,
                    void* acc_maps, void* acc_cube, void* stream);

/* ---- un-fused binning of already projected points (grid.py:70-262) ------------------------- */
int pnm_bin_points(const pnm_grid* g, int64_t n, int dtype,
                   const void* x, const void* y, const void* d, const void* w,
                   const uint8_t* mask, int mask_mode,
                   int sums, int acc_mode, const double* scales3_host,
                   void* acc_maps, void* acc_cube, void* stream);

/* ---- finalisation (grid.py:133-138, :203-205) -----------------------------------------------
 * Reads acc_maps (and acc_cube when `sums` has PNM_W_FROM_CUBE), writes planar float64
 * (ny, nx) maps.  Any output pointer may be NULL.
 *   vmaps: M = S0, V = S1/S0, S = sqrt(S2/S0 - V*V) where S0 != 0, else 0
 *   map  : W = S0, D = S1/S0 where S0 != 0, DW = S1                                            */
int pnm_finalize_vmaps(const pnm_grid* g, const void* acc_maps, const void* acc_cube,
                       int sums, int acc_mode, const double* scales3_host,
                       double* M, double* V, double* S, double* S1, double* S2, void* stream);
int pnm_finalize_map(const pnm_grid* g, const void* acc_maps, int acc_mode,
                     const double* scales3_host, double* W, double* D, double* DW, void* stream);
/* cube accumulator -> float64 cube[ix][iy][iv] (the layout of np.histogramdd, grid.py:259); not in place */
int pnm_finalize_cube(const pnm_grid* g, const void* acc_cube, int sums, int acc_mode, double scale_w,
                      double* cube_out, void* stream);

/* ---- a5: separable Gaussian smoothing of every velocity slice (losvd.py:127-135) ------------
 * cube_in/out/work: float64 [nx][ny][nv], three distinct buffers (work = scratch of the same
 * size, owned by the caller).  taps0_host act along array axis 0, taps1_host along axis 1 (odd
 * counts, already normalised, built on the host in float64).  Zero fill outside the array
 * (astropy convolve boundary='fill', fill_value=0).  NaN voxels follow astropy's default
 * nan_treatment='interpolate': an output with NaNs in its footprint is the kernel-weighted mean of
 * the non-NaN samples (samples outside the array count as 0), or the input value if all are NaN. */
int pnm_smooth_cube(const double* cube_in, double* cube_out, double* work,
                    int32_t nx, int32_t ny, int32_t nv,
                    const double* taps0_host, int32_t ntaps0,
                    const double* taps1_host, int32_t ntaps1, void* stream);

/* ---- f3: integer-factor spatial rebin of every velocity slice (losvd.py:69-92, misc_io.py:383-394)
 * cube_out[i][j][k] = sum (mean != 0: mean) of the factor_x * factor_y block of cube_in, summed
 * first along Y then along X like rebin_2darray.  cube_in (nx, ny, nv) must have nx, ny divisible
 * by the factors (the reference's reshape raises otherwise; the host layer checks).            */
int pnm_rebin_cube(const double* cube_in, double* cube_out, int32_t nx, int32_t ny, int32_t nv,
                   int32_t factor_x, int32_t factor_y, int mean, void* stream);

/* ---- f2: SSP mass-to-light weights (misc_functions.py:195-202) ------------------------------
 * out[i] = scipy.interpolate.griddata(nodes, values, (qx[i], qy[i]), method='cubic'), i.e. the
 * Clough-Tocher C1 cubic on the Delaunay triangulation of the table, NaN outside the convex hull,
 * replaced by the value of the nearest node when fill_nearest != 0.  The triangulation, scipy's
 * gradient estimate and the per-triangle Bezier ordinates are table-only work done once on the
 * host (pynmap_b200/ssp.py); all table pointers are DEVICE pointers:
 *   tri_xform [ntri][6]   T00 T01 T10 T11 r0 r1   (b01 = T (q - r), b2 = 1 - b0 - b1)
 *   tri_coef  [ntri][19]  c3000 c2100 c2010 c2001 c1200 c1101 c1020 c1011 c1002 c0300 c0210 c0201
 *                         c0120 c0111 c0102 c0030 c0021 c0012 c0003
 *   cell_start [ncell*ncell + 1], cell_tris: CSR lists of candidate triangles per cell of a uniform
 *   grid over the table's bounding box (cell = (floor((qx-lo0)*inv0), floor((qy-lo1)*inv1)))
 *   nodes [nnode][3] x y value.   qx, qy: float32 or float64 (dtype); out: float64.                */
int pnm_ct_interp(const void* qx, const void* qy, int64_t n, int dtype,
                  const double* tri_xform, const double* tri_coef, int32_t ntri,
                  const int32_t* cell_start, const int32_t* cell_tris, int32_t ncell,
                  double lo0, double lo1, double inv0, double inv1,
                  const double* nodes, int32_t nnode, int fill_nearest, double* out, void* stream);

/* ---- f1: per-spaxel moments of the cube (losvd.py:41-51, misc_functions.py:70-90) ----------
 * mom0 = sum_k c, mom1 = sum_k v c / mom0, mom2 = sqrt(sum_k v^2 c / mom0 - mom1^2); (nx, ny).
 * An all-zero spaxel gives (0, 0, empty_m2), empty_m2 = min(diff(binv)) / 3 from the host.     */
int pnm_cube_moments(const double* cube, int32_t nx, int32_t ny, int32_t nv,
                     const double* binv_dev, double empty_m2,
                     double* mom0, double* mom1, double* mom2, void* stream);

/* ---- f3: spread_amr (misc_io.py:184-224), used by snapshot.amr_velocity_moments (pynmap.py:342-419)
 * Every AMR cell is replicated nsample times inside its footprint.  The uniform deviates are the
 * CALLER's (the reference draws them from numpy's global generator, misc_io.py:209; the host layer
 * does the same, so a seeded run reproduces the reference), scale = box / 2**level per cell.
 *   pnm_amr_jitter: out[k] = coord[k / nsample] + ((uniform[k] - 0.5) * stretch) * scale[k / nsample], float64
 *   pnm_amr_repeat: out[k] = values[k / nsample] (/ nsample when divide), in the input dtype        */
int pnm_amr_jitter(const void* coord, int dtype, const double* scale_dev, const double* uniform_dev, int64_t n,
                   int32_t nsample, double stretch, double* out, void* stream);
int pnm_amr_repeat(const void* values, int dtype, int64_t n, int32_t nsample, int divide, void* out, void* stream);

/* ---- radial profiles of the velocity ellipsoid (grid.py:16-68 point_2vprofiles; pynmap.py:457-474)
 * Particles with mask != 0 (mask may be NULL) and z_lo < z < z_hi; R = sqrt(x^2 + y^2) rounded in the
 * particle dtype; bin = np.digitize(R, rbin) over nbins increasing float64 radii (bin 0: R < rbin[0];
 * R >= rbin[nbins-1] dropped, as the reference's loop does); VR = Vx cos + Vy sin, VT = -Vx sin + Vy cos
 * with cos = x / R, sin = y / R.  v, s: float32 [3][nbins] = mean and std (ddof 0) of (VR, VT, Vz),
 * NaN for an empty bin; counts: int64 [nbins] or NULL.  acc: workspace of pnm_profile_acc_elems(nbins)
 * float64 (cleared by the call).  pnm_profile_rmax: max R of the selected particles -> out_dev[0]
 * (float64), the reference's default outer radius (grid.py:51).                                   */
int64_t pnm_profile_acc_elems(int32_t nbins);
int pnm_profile_rmax(const void* x, const void* y, const void* z, const uint8_t* mask, int64_t n, int dtype,
                     double z_lo, double z_hi, double* out_dev, void* stream);
int pnm_radial_profiles(const void* x, const void* y, const void* z, const void* vx, const void* vy, const void* vz,
                        const uint8_t* mask, int64_t n, int dtype, double z_lo, double z_hi, const double* rbin_dev,
                        int32_t nbins, double* acc, float* v, float* s, int64_t* counts, void* stream);

/* ---- f4: Gauss-Hermite fit of every spaxel (losvd.py:53-67, fit_losvd.py:57-235) -------------
 * For each of the nspax profiles cube[s][0..nv) minimise sum(((d - A u(p)) / err)^2) over
 * p = (V, S, h3 .. h_deg_gh) with A = sum((d/err)^2) / sum((d/err)(u/err)) re-solved at every
 * evaluation (fit_losvd.py:25-54, :178-185), u = gausshermite(binv, [1, p]) (misc_functions.py:
 * 15-68).  Bounded Levenberg-Marquardt with an analytic Jacobian; limits, convergence tests (ftol,
 * xtol, gtol, maxiter) and status codes as in utils/mpfit.py:1074-1094, :1171-1230, :1303-1322.
 * HOST arrays of deg_gh entries (the parinfo of fit_losvd.py:203-205):
 *   start  start value, NaN = the reference's default: the profile's moments for V and S
 *          (fit_losvd.py:138, :158), guess_h for h_n; may be NULL (all defaults)
 *   lo, hi limits, -inf / +inf where not limited
 *   fixed  non-zero = held at its start value; may be NULL
 * err_dev may be NULL (unit errors).  empty_m2 = min(diff(binv)) / 3 (moment012 of an empty profile).
 *   gh      [deg_gh + 1][nspax]  (I, V, S, h3 ..)      bestfit [nspax][nv] or NULL
 *   status  [nspax]  1..4 converged, 5 maxiter, -16 non-finite (e.g. all-zero profile: I = NaN and
 *                    p = start values, as the reference returns), 0 start values outside the limits
 *                    (the reference raises; gh = NaN)
 *   niter   [nspax]  accepted steps + 1 (mpfit's niter)       chi2 [nspax] final sum of squares
 * deg_gh in 2..6.                                                                                */
int pnm_fit_gh(const double* cube, const double* err_dev, int64_t nspax, int32_t nv, const double* binv_dev,
               int32_t deg_gh, const double* start, const double* lo, const double* hi, const int32_t* fixed,
               double guess_h, double empty_m2, double ftol, double xtol, double gtol, int32_t maxiter,
               double* gh, double* bestfit, int32_t* status, int32_t* niter, double* chi2, void* stream);

/* ---- slab-sharded finalisation (multi-GPU; new: the reference is a single process) ----------
 * The PNM_CUBE_MOMENTS accumulator is an array of record slabs [g][ix][iy][4], g <= G = ceil(nv / 2).  After a
 * reduce-scatter over ranks along g, a rank holds slabs [g0, g1) of the total (`slabs` points at slab g0; slabs past
 * G are padding).  pnm_slab_sums: that rank's share of (sum w, sum w*d, sum w*d*d) per pixel, accumulator type,
 * [3][nx*ny] in (ix, iy) order -- summed over ranks by an all-reduce (exact for int64); pnm_finalize_sums: the totals
 * -> M, V, S (grid.py:133-138, (nY, nX) float64 like pnm_finalize_vmaps); pnm_finalize_cube_slabs: the rank's
 * velocity bins [2 g0, min(nv, 2 g1)) of the LOSVD cube as float64 (nX, nY, bins) (grid.py:259).                 */
/* out[i] = sum over r < nchunks of chunks[r * per + i], added in index order (8-byte accumulator elements, 16-byte
 * aligned arrays): the local reduction of a peer-to-peer reduce-scatter -- every rank pulls its chunk of every
 * rank's partial grid into `chunks` with the copy engines and sums them here, in rank order.                       */
/* cudaMemcpyAsync(dst, src, nbytes, cudaMemcpyDefault, stream): device pointers of this or of a peer device (unified
 * addressing) -- the copy-engine pull of the peer-to-peer reduce-scatter.                                          */
int pnm_copy_async(void* dst, const void* src, int64_t nbytes, void* stream);
int pnm_zero_async(void* dst, int64_t nbytes, void* stream);
/* the pulls of a peer-to-peer reduce-scatter in one call: for every rank p, chunk `rank` (chunk_bytes bytes) of the
 * buffer at bases_host[p] (this device's or a peer's) -> stage + p * chunk_bytes; own buffer first, then ring order.  */
int pnm_pull_chunks(void* stage, const void* const* bases_host, int32_t nranks, int32_t rank, int64_t chunk_bytes, void* stream);      /* cudaMemsetAsync(dst, 0, nbytes, stream) */
/* out[i] = sum over r < nptrs of ptrs_host[r][i] (8-byte accumulator elements; index order): the arrays may live on peer
 * devices mapped into this process (symmetric memory) -- a one-shot all-reduce of small arrays over NVLink loads.    */
int pnm_sum_peers(const void* const* ptrs_host, int32_t nptrs, int64_t n, int acc_mode, void* out, void* stream);
int pnm_sum_chunks(const void* chunks, int32_t nchunks, int64_t per, int acc_mode, void* out, void* stream);
int pnm_slab_sums(const pnm_grid* g, const void* slabs, int32_t g0, int32_t g1, int acc_mode, void* sums3, void* stream);
int pnm_finalize_sums(const pnm_grid* g, const void* sums3, int acc_mode, const double* scales3_host,
                      double* M, double* V, double* S, void* stream);
int pnm_finalize_cube_slabs(const pnm_grid* g, const void* slabs, int32_t g0, int32_t g1, int acc_mode, double scale_w,
                            double* cube_out, void* stream);

/* ---- is the snapshot stored in spatial order?  nsamples pairs of consecutive particles (i, i + 1), i = k * stride:
 * out2_dev[0] = pairs closer than `limit`, out2_dev[1] = pairs examined (device, uint64).  The host layer picks the
 * accumulator layout from it (ordered snapshots merge consecutive particles in registers: pnm_project_bin). ------- */
int pnm_order_probe(const void* x, const void* y, const void* z, int64_t n, int dtype, int64_t stride, int64_t nsamples,
                    double limit, uint64_t* out2_dev, void* stream);

/* ---- max |a[i]| over n elements -> out_dev[0] (float64, device); used to pick FIXED scales -- */
int pnm_absmax(const void* a, int64_t n, int dtype, double* out_dev, void* stream);

/* ---- privatised scatter: pnm_project_bin(sums = W|WD|WD2|CUBE|W_FROM_CUBE|CUBE_MOMENTS) for float32 particles, one
 * view, one matrix, no mask (rotation.py:115-117 + grid.py:126-131 + grid.py:252-259 in one read of the particles).
 * Same accumulator (pnm_cube_moments_acc_elems elements, PNM_CUBE_MOMENTS layout; the maps accumulator is not used),
 * same finalisers.  Every CTA first accumulates, in shared memory, sum w*d and sum w*d*d of the pixels of a box of at
 * most 64 x 64 pixels and sum w of the most populated voxels of that box; what to keep there is decided by a PLAN built
 * on the device from a sample of the particles (pnm_cube_plan_build; rebuild it when the snapshot, the view or the
 * grid changes -- a stale plan only costs speed, never correctness).
 *   PNM_ACC_FIXED: bit-identical to pnm_project_bin with the same scales for any plan.  scales3_host must be powers
 *                  of two in [2^-120, 2^120]; a term stays in shared memory only while |term * 2^k| < 2^31, so scales
 *                  that put the typical term near 2^24 (pnm_cube_plan_info reports such scales) are the fast ones.
 *   PNM_ACC_F64  : privatised terms are rounded to 2^-25 of the sample mean of |term|; particles with a term below
 *                  ~1/16 of the typical w or w*d, above 64 x the typical one or non-finite take the exact RED path,
 *                  so every term is rounded by less than 4.8e-7 of itself: each per-voxel / per-pixel sum agrees
 *                  with pnm_project_bin to better than 4.8e-7 x sum |term|.
 * Particle arrays must be 16-byte aligned; n < 2^32.  Launches that share a plan must be ordered on one stream (the
 * plan owns the kernel's work counter).
 *   pnm_cube_plan_info (synchronises `stream`): out8_host = { box x0, y0, width, height, voxel slots used, threshold
 *   level, sample particles captured, sample size }; stats6_host = sample means of |w|, |w*d|, |w*d*d| and the 2^k
 *   the float64 mode applies to them.                                                                        */
typedef struct pnm_cube_plan pnm_cube_plan;
int32_t pnm_cube_slots_max(void);
int pnm_cube_plan_create(const pnm_grid* g, pnm_cube_plan** out);
int pnm_cube_plan_destroy(pnm_cube_plan* plan);
int pnm_cube_plan_build(const pnm_grid* g, pnm_cube_plan* plan, int64_t n, const float* x, const float* y, const float* z,
                        const float* vx, const float* vy, const float* vz, const float* w, const float* mat9_host,
                        void* stream);
int pnm_cube_plan_info(const pnm_cube_plan* plan, int32_t* out8_host, double* stats6_host, void* stream);
/* SMs pnm_project_cube leaves idle (0 .. 64; default 0): the kernel is persistent with one CTA per SM, so a collective
 * or post-processing kernel of another stream can only run next to it on SMs it does not take (sharded runs).      */
int pnm_cube_plan_set_reserved_sms(pnm_cube_plan* plan, int32_t nsm);
int pnm_project_cube(const pnm_grid* g, const pnm_cube_plan* plan, int64_t n, const float* x, const float* y,
                     const float* z, const float* vx, const float* vy, const float* vz, const float* w,
                     const float* mat9_host, int acc_mode, const double* scales3_host, void* acc_cube, void* stream);

/* ---- host-buffer path: what a reference user calls, end to end -------------------------------
 * Particle arrays live in HOST memory (pinned or pageable).  pnm_accumulate_host streams them
 * through an internal device staging ring in chunks (H2D copy of chunk c+1 overlapping the
 * pnm_project_bin kernel of chunk c) and adds into the caller's DEVICE accumulators; `stream`
 * is made to wait for the binning, and the call returns once the host arrays have been consumed
 * (they may be reused).  One view (nviews = 1).  device = CUDA ordinal.                        */
int pnm_accumulate_host(const pnm_grid* g, int device, int64_t n, int dtype,
                        const void* x_host, const void* y_host, const void* z_host,
                        const void* vx_host, const void* vy_host, const void* vz_host,
                        const void* w_host,
                        const float* mats_host, int nchain,
                        int sums, int acc_mode, const double* scales3_host,
                        void* acc_maps, void* acc_cube, void* stream);

/* pnm_project_host = grid + pnm_accumulate_host + finalise (+ smoothing of the cube when taps are
 * given) + copy of the results to HOST buffers (any may be NULL):
 *   M, V, S : float64 (ny, nx);   cube : float64 (nx, ny, nv).
 * i.e. snapshot.rotate -> velocity_moments -> create_losvds -> LOSVD.convolve_spatial
 * (pynmap.py:246-285, :319-340, :439-455, losvd.py:94-141) in one synchronous call.            */
int pnm_project_host(int device, int64_t n, int dtype,
                     const void* x_host, const void* y_host, const void* z_host,
                     const void* vx_host, const void* vy_host, const void* vz_host,
                     const void* w_host,
                     const float* mats_host, int nchain,
                     int32_t nx, const double* edges_x_host, int32_t ny, const double* edges_y_host,
                     int32_t nv, const double* edges_v_host,
                     int acc_mode, const double* scales3_host,
                     const double* taps0_host, int32_t ntaps0, const double* taps1_host, int32_t ntaps1,
                     double* M_host, double* V_host, double* S_host, double* cube_host);
/* frees the staging ring pnm_project_host keeps between calls */
int pnm_host_release(void);

#ifdef __cplusplus
}
#endif
#endif /* PNM_B200_H */
