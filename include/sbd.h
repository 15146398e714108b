/*
 * sbd.h -- C-ABI of the B200-native sea-breeze trigger ("sbd" = sea-breeze diagnostic).
 *
 * Drop-in boundary for the hot path of antarcticrainforest/seabreeze_param.  Every entry point
 * replaces one reference routine; citations are file:line under the reference tree:
 *   G = generic/sea_breeze_diag.f90             (module sea_breeze_diag_mod, Fortran host models)
 *   H = generic/halo_exchange_mod.f90           (swap_bounds)
 *   P = python_wrapper/seabreezediag/seabreeze_diag_python.f90   (f2py module `seabreeze`)
 *   S = python_wrapper/seabreezediag/sobel.f90                  (f2py module `seabreeze`)
 *   U = UM/vn10.7/sea_breeze_diag.F90           (UM hook; same C-ABI, `error` = return status)
 *
 * Conventions
 *  - plain C, no CUDA / torch types.  All arrays are binary32, column-major (Fortran order),
 *    longitude (x) fastest: a(j,i[,k]) -> a[(k*ny + i)*nx + j] with 0-based i,j,k.
 *  - every pointer may be a HOST pointer (pageable or pinned) or a DEVICE pointer on the current
 *    CUDA device; the library detects which (cudaPointerGetAttributes).  Host data is staged
 *    through library-owned device buffers and the call blocks until outputs are back (reference
 *    semantics).  With device pointers the work is enqueued on the plan's stream and the call
 *    returns immediately (use sbd_plan_sync / the stream to wait).
 *  - return value: 0 = ok, 1 = bad dimensions (the UM variant's `error = 1`, U:198-201),
 *    2 = invalid argument, 3 = CUDA failure (sbd_last_error() has the text), 4 = a neighbour rank did not
 *    deliver its halo rows in time (multi-GPU row blocks).
 *  - there is NO CPU fallback: without a usable CUDA device every entry point returns 3.
 *  - transcendental functions are the pinned algorithms of sbd_math.h (see there).
 *
 * "halo'd" arrays are (nx+2h) x (ny+2h) with the nx x ny interior at offset (h,h)
 * (UM convention tdims_s / tdims_l, U:78-117).
 */
#ifndef SBD_H
#define SBD_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SBD_OK 0
#define SBD_EDIMS 1
#define SBD_EINVAL 2
#define SBD_ECUDA 3
#define SBD_EPEER 4 /* multi-GPU: a neighbour rank never posted its halo rows (sbd_plan_set_peer_sync) */

#define SBD_VARIANT_GENERIC 0 /* G / U: halo'd neighbour fields, 3-D p, windspeed updated every step */
#define SBD_VARIANT_PYTHON 1  /* P / S: index-mapped boundaries, 1-D p, 6-hourly windspeed update    */

/* Tunables.  Generic/UM: compile-time parameters G:127-138 / U:158-169.  Python: runtime optional
 * arguments P:137-141 (already converted to SI here: Pa, s). */
typedef struct sbd_params {
  float target_plev;    /* Pa   (70000)     */
  float thresh_wind;    /* m/s  (11)        */
  float thresh_winddir; /* deg  (90)        */
  float thresh_windch;  /* m/s  (5)         */
  float thresh_thc;     /* K    (0.75)      */
  float target_time;    /* s    (21600)     */
  float maxdist;        /* km   (180)       */
  float gmma;           /* K/m  (-0.0060956)*/
  float rad2deg;        /*      (57.2957)   */
  float missing;        /* sb_con far from the coast: 0.0 (G:176) or 2.0E20 (P:173) */
  int level_scan;       /* 0 = minloc (G:223), 1 = early-exit scan (U:265-274) */
} sbd_params;

void sbd_default_params(sbd_params* prm, int variant);
const char* sbd_last_error(void);
/* Number of usable CUDA devices (0 => every compute entry point fails with SBD_ECUDA). */
int sbd_device_count(void);

/* ---------------------------------------------------------------------------------------------
 * Generic (Fortran host model) entry points: what an ISO_C_BINDING shim named
 * sea_breeze_diag_mod forwards to (fortran/sea_breeze_diag_mod.f90).
 * ------------------------------------------------------------------------------------------- */

/* get_edges(mask, icefrac, landfrac, halo_size)  G:273-373.
 * mask: halo'd, out.  icefrac, landfrac: nx x ny.  swap != 0 fills the halos at the two
 * swap_bounds call sites (G:342, G:371) with sbd_swap_bounds_f32; swap == 0 leaves the halos as
 * the caller supplied them (the literal no-op stub H:12-17).
 * m1 (optional, nx*ny bytes): binarised land/ice mask (G:325-337). */
int sbd_get_edges_f32(float* mask, const float* icefrac, const float* landfrac, int nx, int ny, int h,
                      int swap, uint8_t* m1);
/* get_edges plus two optional by-products of its Sobel stencil: px, py
 * (nx*ny each, may be NULL) receive the Sobel components of generic/sea_breeze_diag.f90:350-357;
 * atan2(py, px) is the direction of the coastline normal.  Not outputs of the reference routine. */
int sbd_get_edges_ex_f32(float* mask, const float* ice, const float* land, int nx, int ny, int h, int swap,
                         uint8_t* m1, float* px, float* py);

/* get_dist(coast, landfrac, lon, lat, maxdist, cdist, halo_size)  G:375-455.
 * coast: halo'd output of get_edges; landfrac, cdist: nx x ny; lon(nx), lat(ny) in degrees.
 * Search radius = h cells (G:422,425). */
int sbd_get_dist_f32(const float* coast, const float* landfrac, const float* lon, const float* lat,
                     float maxdist, float* cdist, int nx, int ny, int h);

/* sigmoid(ary, sm)  G:457-481.  stats (optional, 4 floats): mean, 2/sigma, r, var. */
int sbd_sigmoid_f32(const float* ary, float* sm, int nx, int ny, float* stats);

/* seabreeze_diag(timestep, timestep_number, p,u,v, theta, mask, z, sigma, windspeed, winddir, thc,
 * sb_con)  G:55-271 (U:55-326 with level_scan=1; its `error` is the return value).
 * p,u,v: nx x ny x nz.  theta, mask (signed coast distance), z, sigma: halo'd.
 * windspeed, winddir, thc, sb_con: nx x ny, inout.  flags (optional, nx*ny bytes):
 * bit0 = coastal zone |mask|<=maxdist (G:174), bit1 = trigger (G:249-252). */
int sbd_seabreeze_diag_f32(float timestep, int timestep_number, const float* p, const float* u,
                           const float* v, const float* theta, const float* mask, const float* z,
                           const float* sigma, float* windspeed, float* winddir, float* thc,
                           float* sb_con, int nx, int ny, int nz, int h, const sbd_params* prm,
                           uint8_t* flags);

/* The one-shot entry points above (and sbd_py_diag_f32 below) keep a small process-wide cache of plans keyed by
 * variant, dimensions and tunables, so that the call a host model makes every timestep
 * (generic/dummy_model.f90:52-54) does not redo the static work: the three static planes are uploaded (or read
 * in place when they are device pointers), checksummed on the device, and the sigmoid statistics / search radii /
 * work lists are rebuilt only when their content changed.  sbd_plan_cache_clear frees the cached plans. */
int sbd_plan_cache_clear(void);
int sbd_plan_cache_stats(int* hits, int* misses, int* plans);

/* Page-lock and device-map a caller's host array (cudaHostRegister): later calls that receive it move data at
 * PCIe rate, and p, u, v are gathered in place by the level scan (only the columns of coastal-zone cells cross the
 * bus).  The range must stay allocated until sbd_host_unregister.  A Fortran host registers its model arrays once
 * at start-up; seabreeze_param_b200 registers numpy arrays with a finalizer (SBD_AUTO_REGISTER=0 disables). */
int sbd_host_register(const void* ptr, size_t bytes);
int sbd_host_unregister(const void* ptr);

/* swap_bounds(field, halo_size)  H:12-17 (stub) / U:408-410.  Single-domain halo fill:
 * east-west periodic, north-south edge-replicated.  The multi-GPU halo swap (NCCL send/recv between
 * row blocks) lives in seabreeze_param_b200/halo.py and uses sbd_swap_bounds_ew_f32 for the local
 * periodic part. */
int sbd_swap_bounds_f32(float* field, int nx, int ny, int h);

/* ---- 64-bit REAL twins ------------------------------------------
 * For hosts compiled with promoted reals (the UM's usual -r8 build: real(kind=jprb) at
 * UM/vn10.7/sea_breeze_diag.F90:64).  Same arguments as the _f32 routines with double arrays and
 * scalars (host or device pointers).
 *   sbd_seabreeze_diag_f64, sbd_sigmoid_f64 (and sbd_um_seabreeze_diag_f64): evaluated in BINARY64 THROUGHOUT --
 *     sequential statistics, t0, window sums in the reference's raster order, level search, wind, thresholds,
 *     scaling, state (csrc/sbd_f64.cuh; U:144-150 accumulate in jprb).  Plain one-thread-per-cell kernels: this is
 *     the correctness path for -r8 hosts, not the tuned fp32 path.  Against a float64 evaluation the results agree
 *     to ~1e-12 relative (CUDA's binary64 exp / atan2 vs glibc's); tests/test_gpu_f64.py states the tolerances.
 *   sbd_get_edges_f64, sbd_get_dist_f64, sbd_swap_bounds_f64 (static, one-off work): the arrays are narrowed to
 *     fp32 on the device, the fp32 kernels run, the results are widened exactly:
 *     twin(x) == widen(f32 routine(narrow(x))) bit for bit; coast distances agree with a 64-bit evaluation to
 *     fp32 accuracy. */
int sbd_get_edges_f64(double* mask, const double* icefrac, const double* landfrac, int nx, int ny, int h,
                      int swap);
int sbd_get_dist_f64(const double* coast, const double* landfrac, const double* lon, const double* lat,
                     double maxdist, double* cdist, int nx, int ny, int h);
int sbd_sigmoid_f64(const double* ary, double* sm, int nx, int ny);
int sbd_seabreeze_diag_f64(double timestep, int timestep_number, const double* p, const double* u,
                           const double* v, const double* theta, const double* mask, const double* z,
                           const double* sigma, double* windspeed, double* winddir, double* thc,
                           double* sb_con, int nx, int ny, int nz, int h, const sbd_params* prm,
                           uint8_t* flags);
int sbd_swap_bounds_f64(double* field, int nx, int ny, int h);
int sbd_swap_bounds_ew_f32(float* field, int nx, int ny, int h);
/* Asynchronous device-only form: one kernel on `stream` (cudaStream_t as void*).  ns_mask bit0 /
 * bit1: replicate the top / bottom edge rows (clear the bit for a side owned by a neighbour rank). */
int sbd_swap_bounds_dev_f32(float* d_field, int nx, int ny, int h, int ns_mask, void* stream);

/* ---------------------------------------------------------------------------------------------
 * UM hook (UM/vn10.7/sea_breeze_diag.F90 = U): the UM's own dummy lists (fortran/sea_breeze_diag_um.F90).
 * Planes are row_length x rows interiors with the halo widths of atm_fields_bounds_mod:
 *   tdims_s (halo_s on every side): theta, z, sigma (U:78-117);   tdims_l (halo_i x halo_j): mask / coast /
 *   landfrac (U:339-341, U:471-476);   pdims / tdims (no halo): p, u, v, windspeed, winddir, thc, sb_con.
 * Return value = the routine's `error` (U:102, U:198-201: 1 for rows < 1 or levels < 1).
 * ------------------------------------------------------------------------------------------- */

/* seabreeze_diag(timestep, timestep_number, p, u, v, theta, z, sigma, mask, windspeed, winddir, thc, sb_con, error)
 * U:55-326.  mask (signed coast distance) has halo_l, theta / z / sigma halo_s; the growing window (U:234-258) may
 * reach min(halo_s, halo_l) cells.  Level search: U's early-exit scan (U:265-274) unless prm says otherwise.
 * theta_to_t0 != 0 reproduces the side effect of U:210-211: theta (the whole halo'd array) is overwritten by
 * t0 = theta - gmma z sigmoid(sigma). */
int sbd_um_seabreeze_diag_f32(float timestep, int timestep_number, const float* p, const float* u, const float* v,
                              float* theta, const float* z, const float* sigma, const float* mask, float* windspeed,
                              float* winddir, float* thc, float* sb_con, int row_length, int rows, int levels,
                              int halo_s, int halo_l, int theta_to_t0, const sbd_params* prm, uint8_t* flags);
/* get_edges(mask, icefrac, landfrac)  U:328-446.  mask: tdims_l (out: coast flag + swapped halos). */
int sbd_um_get_edges_f32(float* mask, const float* icefrac, const float* landfrac, int row_length, int rows,
                         int halo_i, int halo_j);
/* get_dist(landfrac, coast)  U:448-601.  landfrac, coast: tdims_l (coast in: flag, out: signed distance + swapped
 * halos); true_latitude / true_longitude: row_length x rows, degrees (trignometric_mod, U:461); search radius
 * halo_i x halo_j cells (U:515-516); maxdist = 180 in U (module parameter). */
int sbd_um_get_dist_f32(const float* landfrac, float* coast, const float* true_latitude, const float* true_longitude,
                        int row_length, int rows, int halo_i, int halo_j, float maxdist);
/* swap_bounds(field, row_len, rows, levels, halo_i, halo_j, fld_type, l_vector)  U:408-410 (halo_exchange module of
 * the UM), single domain: east-west periodic, north-south edge rows.  fld_type / l_vector do not matter for scalars
 * on one domain; between GPUs see seabreeze_param_b200/halo.py. */
int sbd_um_swap_bounds_f32(float* field, int row_len, int rows, int levels, int halo_i, int halo_j);

/* 64-bit REAL twins of the four UM entry points (fortran/sea_breeze_diag_um.F90 compiled with -DSBD_R8):
 * seabreeze_diag in binary64 throughout, the other three as described for the _f64 twins above. */
int sbd_um_seabreeze_diag_f64(double timestep, int timestep_number, const double* p, const double* u, const double* v,
                              double* theta, const double* z, const double* sigma, const double* mask, double* windspeed,
                              double* winddir, double* thc, double* sb_con, int row_length, int rows, int levels,
                              int halo_s, int halo_l, int theta_to_t0, const sbd_params* prm, uint8_t* flags);
int sbd_um_get_edges_f64(double* mask, const double* icefrac, const double* landfrac, int row_length, int rows,
                         int halo_i, int halo_j);
int sbd_um_get_dist_f64(const double* landfrac, double* coast, const double* true_latitude, const double* true_longitude,
                        int row_length, int rows, int halo_i, int halo_j, double maxdist);
int sbd_um_swap_bounds_f64(double* field, int row_len, int rows, int levels, int halo_i, int halo_j);

/* ---------------------------------------------------------------------------------------------
 * Python (f2py module `seabreeze`) entry points: same positional meaning as the f2py signatures.
 * ------------------------------------------------------------------------------------------- */

/* coast = get_edges(lsm, ci, [nlons, nlats])  S:19-89 */
int sbd_py_get_edges_f32(const float* lsm, const float* ci, int nlons, int nlats, float* coast);

/* cdist = get_dist(coast, mask, lon, lat, [nlons, nlats, maxdist=180])  S:91-193.
 * k_out (optional): the search radius derived at S:129-137. */
int sbd_py_get_dist_f32(const float* coast, const float* mask, const float* lon, const float* lat,
                        int nlons, int nlats, float maxdist, float* cdist, int* k_out);

/* sm = sigmoid(ary, [nlons, nlats])  P:287-311 */
int sbd_py_sigmoid_f32(const float* ary, int nlons, int nlats, float* sm);

/* output = diag(timestep_number, p, z, std, theta, v, u, cdist, windspeed, winddir, thc,
 *               [target_plev=700, thresh_wind=11, thresh_winddir=90, thresh_windch=5,
 *                thresh_thc=0.75, target_time=6, maxdist=180, timestep=24, nps, nlons, nlats])
 * P:49-285.  Units as in the f2py signature (hPa, h, min).  output: (nlons,nlats,4) =
 * sb_con, t0, windspeed, winddir.  flags optional as above. */
int sbd_py_diag_f32(int timestep_number, const float* p, const float* z, const float* std,
                    const float* theta, const float* v, const float* u, const float* cdist,
                    const float* windspeed, const float* winddir, const float* thc, float target_plev,
                    float thresh_wind, float thresh_winddir, float thresh_windch, float thresh_thc,
                    float target_time, float maxdist, float timestep, int nps, int nlons, int nlats,
                    float* output, uint8_t* flags);

/* The whole time loop of seabreezediag.diag() (python_wrapper/seabreezediag/__init__.py:219-245) in
 * one call: for r = 0..nrec-1: dist_r = get_dist(get_edges(lsm, ci_r), lsm, lon, lat);
 * output_r = diag(timestep_number0 + r, p, z, std, theta_r, v_r, u_r, dist_r, ws_{r-1}, wd_{r-1}, ...)
 * with ws_r, wd_r = planes 3,4 of output_r (I:244) and ws_{-1}, wd_{-1} = ws0, wd0.
 * theta (nlons,nlats,nrec); u,v (nlons,nlats,nps,nrec); ci (nlons,nlats,nrec) or NULL (no ice);
 * output (nlons,nlats,4,nrec); flags optional (nlons,nlats,nrec).  Everything stays on the device
 * between records; the coast distance is recomputed only when ci_r differs from ci_{r-1}. */
int sbd_py_diag_loop_f32(int timestep_number0, int nrec, const float* lsm, const float* ci, const float* lon,
                         const float* lat, const float* z, const float* std, const float* p, const float* u,
                         const float* v, const float* theta, const float* ws0, const float* wd0,
                         float target_plev, float thresh_wind, float thresh_winddir, float thresh_windch,
                         float thresh_thc, float target_time, float maxdist, float timestep, int nps, int nlons,
                         int nlats, float* output, uint8_t* flags);

/* nt = get_threads()  S:195-206: here the number of CUDA devices. */
int sbd_get_threads(void);

/* ---------------------------------------------------------------------------------------------
 * Plan API: device-resident, static fields processed once (what a host model that calls
 * seabreeze_diag every timestep with unchanged mask/z/sigma wants).
 * ------------------------------------------------------------------------------------------- */
typedef struct sbd_plan sbd_plan;

/* variant: SBD_VARIANT_*.  nz: levels of p,u,v (python: nps).  h: halo (generic) -- ignored for
 * python.  prm may be NULL (defaults of the variant). */
int sbd_plan_create(sbd_plan** plan, int variant, int nx, int ny, int nz, int h, const sbd_params* prm);
void sbd_plan_destroy(sbd_plan* plan);
/* Use an external CUDA stream (cudaStream_t cast to void*); NULL is the legacy default stream.
 * sbd_plan_use_own_stream restores the plan's own non-blocking stream. */
int sbd_plan_set_stream(sbd_plan* plan, void* stream);
int sbd_plan_use_own_stream(sbd_plan* plan);
void* sbd_plan_get_stream(sbd_plan* plan);
int sbd_plan_sync(sbd_plan* plan);

/* Generic variant: fold the single-domain swap_bounds of the per-step field theta (and of the
 * static planes) into the kernels' reads instead of reading halo memory.  mode is a bit mask:
 * 1 = east-west periodic, 2 = north edge row replicated, 4 = south edge row replicated; 0 (default)
 * = read the halos the caller filled (reference semantics).  A row-block rank sets 1 plus the
 * bit of each side that is a domain edge and exchanges only the other side's rows.  Call before
 * sbd_plan_set_static. */
int sbd_plan_set_halo_mode(sbd_plan* plan, int mode);
/* Multi-GPU row blocks without a per-step copy: device pointers to the halo'd theta planes of the
 * northern / southern neighbour rank, mapped into this process (CUDA peer or symmetric memory over
 * NVLink).  The window kernel then reads the halo rows it needs straight from the neighbours'
 * interior rows; NULL = that side's halo comes from this plane (or the halo mode).  ny_up = interior
 * rows of the northern neighbour.  The pointers stay in force until changed. */
int sbd_plan_set_peer_theta(sbd_plan* plan, const float* theta_up, int ny_up, const float* theta_dn);
/* Per-step synchronisation of the neighbour reads above without a separate barrier launch.
 * sync_local: >= 4 zero-initialised uint32 words in this rank's device memory that the neighbours
 * can write (peer / symmetric memory); sync_up / sync_dn: the same array of the northern / southern
 * neighbour as mapped into this process (NULL where there is no neighbour).  Every step the window
 * kernel posts its step sequence number into the neighbours' words and only the CTAs that read
 * neighbour rows wait for the neighbours' own post.  Every rank must issue the same number of
 * steps; a rank's theta plane of step k must stay untouched until its step k+1 has completed
 * (keep at least two theta planes and alternate).  A neighbour that never posts is reported after ~4 s instead of
 * hanging the GPU: the kernel leaves a marker in sync_local[2], [3], and sbd_plan_sync -- and every step call that
 * synchronises because it was given host arrays -- then returns SBD_EPEER (the step's results are invalid).
 * (Re-)arming clears this rank's four words and restarts the sequence: do it on every rank between two barriers.
 * NULL sync_local switches the handshake off. */
int sbd_plan_set_peer_sync(sbd_plan* plan, uint32_t* sync_local, uint32_t* sync_up, uint32_t* sync_dn);
/* Form of the window-sum kernel: 0 = automatic (tile kernel k_trigger for search radii <= 8, the pair-layout strip
 * kernel k_win4 for 9..31, the signed-class strip kernel k_win8 from 32), 1 = tile kernel (one cell per thread, cells
 * bucketed by radius), 2 = k_win4, 3 = k_win8 (four CTAs per SM up to radius 31, two beyond).  Every form evaluates each cell's sums in the reference's raster
 * order: identical bits (tests force each form on every grid class). */
int sbd_plan_set_window_kernel(sbd_plan* plan, int mode);
/* Cells per strip of k_win8: 4 (default) or 8 (faster only where a tile's radii are nearly uniform). */
int sbd_plan_set_strip_cells(sbd_plan* plan, int cells);
/* k_win8 can give a tile whose window work exceeds a fair share of an SM's time to up to max_parts CTAs (each stages
 * the tile and takes every max_parts-th strip); call before sbd_plan_set_static.  Default 1 = never split: measured on
 * one eighth of the c3 grid (a rank's share at 8 GPUs) the window kernel takes 0.70 / 0.72 / 0.81 / 0.87 / 0.90 ms with
 * 1 / 2 / 3 / 4 / 8 parts -- the duplicated staging and the thinner chunks cost more than the shorter tail saves. */
int sbd_plan_set_tile_split(sbd_plan* plan, int max_parts);
/* Rows per tile of k_win8: 32 (default, also what 0 = auto resolves to) or 16.  Half-height tiles halve the longest
 * CTA of a launch but were measured slower on every grid tried (a 4000 x 384 block of c3: 0.59 ms against 0.48 ms);
 * kept for experiments. */
int sbd_plan_set_tile_rows(sbd_plan* plan, int rows);
/* Strip window kernels (k_win4, k_win8): 1 = a streaming pre-pass per step (k_xplane) writes x = +-t0 of every cell a
 * tile may stage -- index maps, virtual halos and neighbour-rank rows resolved once per cell -- and the tiles stage
 * from it with one coalesced load per cell; 0 (default) = tiles stage from theta / gzsm / code themselves.  Same bits
 * either way (tested).  Measured, window kernel incl. the pre-pass: c3 2.94 ms against 2.95 ms, a c4-sized plane 296 us against
 * 277 us -- the tiles' staging is latency that the SM's other CTA covers, the pre-pass is 15-25 us of its own. */
int sbd_plan_set_xplane(sbd_plan* plan, int on);
/* Process-wide: 1 = evaluate the sigmoid statistics with the plain sequential chain (one warp) instead of the exact
 * parallel formulation.  Same bits either way; kept as the in-library cross-check of the parallel form. */
int sbd_set_stats_mode(int sequential);
/* Restrict the trigger to rows [row_lo, row_hi) of the plan's grid (row_hi < 0: to the end); call before
 * sbd_plan_set_static.  Cells of other rows are treated like cells outside the coastal zone (missing
 * value, state untouched) but still serve as neighbours.  This is how the python variant -- whose
 * index map clamps rows at the array edge -- is sharded over GPUs without any exchange: a rank's plan
 * holds its rows plus an overlap of >= the largest search radius on either side and processes only its
 * own rows (seabreeze_param_b200/halo.py: PyRowBlock).  The python variant's "last row is never
 * processed" (seabreeze_diag_python.f90:165) still applies to the plan's last row. */
int sbd_plan_set_row_window(sbd_plan* plan, int row_lo, int row_hi);
/* Process-wide default for plans left at mode 0 (also used by the one-shot entry points). */
int sbd_set_window_kernel_default(int mode);
/* Static fields: mask = signed coast distance, z, sigma (generic: halo'd; python: nlons x nlats).
 * Computes the sigmoid statistics (sequential-order exact), the sea-level correction plane
 * gmma*z*sigmoid(sigma), the land/sea sign plane and every cell's final search radius nn.
 * stats_in (optional, 3 floats mean, 2/sigma, r): use externally chained global statistics
 * (multi-GPU row blocks) instead of computing them from this block. */
int sbd_plan_set_static(sbd_plan* plan, const float* mask, const float* z, const float* sigma,
                        const float* stats_in);
/* Largest search radius found by set_static and the sigmoid statistics (mean, 2/sigma, r, var). */
int sbd_plan_info(sbd_plan* plan, int* nn_max, int* n_coastal, float* stats4);

/* By-products of the static planes (any pointer may be NULL; int32, nx*ny,
 * host or device): nn = the search radius at which the growing-window loop of
 * generic/sea_breeze_diag.f90:188-214 stops for each cell, n_land / n_sea = its counters n_l, n_s at
 * that pass; all 0 outside the coastal zone. */
int sbd_plan_window_info(sbd_plan* plan, int* nn, int* n_land, int* n_sea);

/* Number of aligned 4-cell groups with at least one coastal-zone cell (the units the level scan
 * streams; 16 bytes per group and level).  -1 before sbd_plan_set_static. */
int sbd_plan_n_quads(sbd_plan* plan);
/* Host-buffer steps with pinned p use a hybrid transfer: the block of rows that holds the northern
 * quads (from the first row with a coastal-zone cell to the split row) is copied level by level by
 * the DMA engine, while the level scan of the remaining quads gathers its columns over PCIe in
 * place.  Returns the number of rows copied and how many quads of the compacted list they serve. */
int sbd_plan_host_split(sbd_plan* plan, int* ysplit, int* n_quads_north);
/* One timestep, generic variant (G:170-269 after G:166-167). */
int sbd_plan_step(sbd_plan* plan, float timestep, int timestep_number, const float* p, const float* u,
                  const float* v, const float* theta, float* windspeed, float* winddir, float* thc,
                  float* sb_con, uint8_t* flags);
/* The same step in two calls, for callers that have something to do in between (multi-GPU: the
 * halo exchange of theta runs while the level scan streams p).  sbd_plan_step_scan needs p,u,v
 * only; sbd_plan_step_finish needs theta (with valid halos) and the state arrays. */
int sbd_plan_step_scan(sbd_plan* plan, const float* p, const float* u, const float* v);
int sbd_plan_step_finish(sbd_plan* plan, float timestep, int timestep_number, const float* theta,
                         float* windspeed, float* winddir, float* thc, float* sb_con, uint8_t* flags);
/* One time record, python variant (P:160-284).  timestep in minutes as in P. */
int sbd_plan_step_py(sbd_plan* plan, int timestep_number, float timestep_min, const float* p,
                     const float* u, const float* v, const float* theta, const float* windspeed,
                     const float* winddir, float* output, uint8_t* flags);
/* Per-kernel CUDA-event timing of the step (off by default), events recorded on the stream the
 * kernels are launched on.  While timing is on the two kernels of the generic step run strictly
 * one after the other (no programmatic dependent launch), so each duration is that kernel alone:
 * ms3[0] = window kernel (k_trigger), ms3[1] = level-scan kernel (k_scan_quad),
 * ms3[2] = threshold kernel (k_final).  When the step runs as the single fused kernel
 * (k_step_fused: coarse grids, device-resident p) and for the python variant: ms3[0] = that one
 * kernel, ms3[1] = ms3[2] = 0. */
int sbd_plan_set_timing(sbd_plan* plan, int on);
int sbd_plan_last_timing(sbd_plan* plan, float* ms3);
/* Number of kernel launches issued by this library since load (for bench.py's gpu_launches). */
unsigned long long sbd_launch_count(void);

/* Sequential-order statistics with carry, for chaining across row blocks in rank order:
 * pass 0: carry[0] = running sum (in/out), carry[2], carry[3] = running max, min (in/out; first=1
 *         initialises them from the data);
 * pass 1: carry[1] = running variance numerator (in/out) given mean in carry[4]. */
int sbd_stats_chain_f32(const float* ary, int nx, int ny, int ld, int pass, int first, float* carry5);

#ifdef __cplusplus
}
#endif
#endif /* SBD_H */
