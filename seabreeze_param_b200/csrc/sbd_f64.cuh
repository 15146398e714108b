// sbd_f64.cuh -- 64-bit arithmetic for hosts compiled with promoted reals (the UM's usual build:
// real(kind=jprb), UM/vn10.7/sea_breeze_diag.F90:64, accumulators U:144-150).  Included by sbd_cuda.cu inside
// namespace sbdk.  Round 1 narrowed such hosts' arrays to fp32, ran the fp32 kernels and widened the results; here
// sigmoid and seabreeze_diag are evaluated in binary64 throughout: sequential statistics, t0, the window sums in the
// reference's raster order, the level search, wind, thresholds, scaling and state.  These kernels are the plain
// statement of the algorithm (one thread per cell, windows read from global memory): correctness for -r8 hosts,
// not the tuned fp32 path.  Transcendentals are CUDA's binary64 exp / atan2 (<= 1-2 ulp from glibc's): results agree
// with a float64 evaluation to ~1e-12 relative, flags can differ only where a quantity sits within that of a
// threshold (tests/test_gpu_f64.py states the tolerances).

// sequential sum / variance / max / min of the interior of a plane with row pitch ld (G:466-479), one warp
__global__ void k64_stats_seq(const double* __restrict__ ary, int nx, int ny, int ld, double* __restrict__ st)
{
  const int lane = threadIdx.x;
  const long long n = (long long)nx * ny;
  double mx = -__longlong_as_double(0x7ff0000000000000LL), mn = __longlong_as_double(0x7ff0000000000000LL);
  double mean = 0.0;
  for (int pass = 0; pass < 2; ++pass) {
    double acc = 0.0;
    for (long long base = 0; base < n; base += 32) {
      const long long t = base + lane;
      double v = 0.0;
      if (t < n) {
        const int y = (int)(t / nx), x = (int)(t % nx);
        v = ary[(size_t)y * ld + x];
        if (pass == 0) { mx = fmax(mx, v); mn = fmin(mn, v); }
        else { const double d = v - mean; v = d * d; }
      }
      const int cnt = (int)min(32LL, n - base);
      for (int k = 0; k < cnt; ++k) acc = acc + __shfl_sync(0xffffffffu, v, k);
    }
    if (pass == 0) {
      mean = acc / (double)n;                                  // G:466
      for (int o = 16; o; o >>= 1) {
        mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        mn = fmin(mn, __shfl_xor_sync(0xffffffffu, mn, o));
      }
    } else if (lane == 0) {
      st[0] = mean;
      st[1] = 2.0 / sqrt(acc / (double)n);                     // G:478
      st[2] = (mx - mn) / 4.0;                                 // G:479
      st[3] = acc;
    }
  }
}

__global__ void k64_sigmoid_apply(const double* __restrict__ ary, double* __restrict__ sm, long long n,
                                  const double* __restrict__ st)
{
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n) return;
  sm[t] = 1.0 / (1.0 + exp((-st[1]) * (ary[t] - st[2])));      // G:480
}

// t0 = theta - (gmma z) sigmoid(sigma) over a halo'd plane (G:167); in_place: theta itself is overwritten (U:210-211)
__global__ void k64_t0(const double* __restrict__ theta, const double* __restrict__ z, const double* __restrict__ sigma,
                       double* __restrict__ t0, long long n, const double* __restrict__ st, double gmma)
{
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n) return;
  const double sm = 1.0 / (1.0 + exp((-st[1]) * (sigma[t] - st[2])));
  t0[t] = theta[t] - ((gmma * z[t]) * sm);
}

struct Diag64Args {
  int nx, ny, nz, h, ldx;
  int first_step, six_hourly, level_scan;
  double target_plev, thresh_wind, thresh_winddir, thresh_windch, thresh_thc, maxdist, rad2deg, missing;
  const double *p, *u, *v, *t0, *mask;
  double *ws, *wd, *thc, *sb;
  uint8_t* flags;
};

__device__ __forceinline__ double modulo64(double a, double p)
{
  double r = fmod(a, p);
  if (r != 0.0 && ((r < 0.0) != (p < 0.0))) r += p;
  return r;
}

// one thread per interior cell: G:170-269 in binary64 (growing window as written: only the last pass's sums
// survive, so the final window is summed once after the radius is known)
__global__ void __launch_bounds__(128) k64_diag(const Diag64Args a)
{
  const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
  if (x >= a.nx) return;
  const size_t c = (size_t)y * a.nx + x;
  const size_t ch = (size_t)(y + a.h) * a.ldx + (x + a.h);
  const size_t n2 = (size_t)a.nx * a.ny;
  const double mk = a.mask[ch];
  if (fabs(mk) > a.maxdist) {                                   // G:174-176
    a.sb[c] = a.missing;
    if (a.flags) a.flags[c] = 0;
    return;
  }
  const double mul = (mk >= 0.0) ? 1.0 : -1.0;                  // G:182-186
  const int nncap = max(a.h, 1);
  int nn = 1;
  for (;;) {                                                     // G:191-214: radius at which both classes are present
    bool hl = false, hs = false;
    for (int ii = y - nn; ii <= y + nn && !(hl && hs); ++ii)
      for (int jj = x - nn; jj <= x + nn; ++jj) {
        if (a.mask[(size_t)(ii + a.h) * a.ldx + (jj + a.h)] >= 0.0) hl = true; else hs = true;
      }
    if ((hl && hs) || nn >= nncap) break;
    nn = nn + 1;
  }
  double T_l = 0.0, T_s = 0.0, n_l = 0.0, n_s = 0.0;
  for (int ii = y - nn; ii <= y + nn; ++ii)
    for (int jj = x - nn; jj <= x + nn; ++jj) {
      const size_t w = (size_t)(ii + a.h) * a.ldx + (jj + a.h);
      if (a.mask[w] >= 0.0) { T_l = T_l + a.t0[w]; n_l = n_l + 1.0; }     // G:200-202
      else { T_s = T_s + a.t0[w]; n_s = n_s + 1.0; }                        // G:204-205
    }
  const double n_thc = mul * ((T_l / n_l) - (T_s / n_s));       // G:216
  int lev = 0;
  if (a.level_scan == 0) {                                       // minloc, first minimum (G:223)
    double best = 0.0;
    for (int k = 0; k < a.nz; ++k) {
      const double d = fabs(a.p[(size_t)k * n2 + c] - a.target_plev);
      if (k == 0 || d < best) { best = d; lev = k; }
    }
  } else {                                                       // U:265-274
    double diff = 1000000.0;
    for (int k = 0; k < a.nz; ++k) {
      const double d = fabs(a.p[(size_t)k * n2 + c] - a.target_plev);
      if (d <= diff) { lev = k; diff = d; } else break;
    }
  }
  const double uu = a.u[(size_t)lev * n2 + c], vv = a.v[(size_t)lev * n2 + c];
  const double n_ws = sqrt(uu * uu + vv * vv);                   // G:225
  const double n_wd = atan2(-1.0 * uu, -1.0 * vv) * a.rad2deg;   // G:227
  double ws = a.ws[c], wd = a.wd[c];
  if (a.first_step) { ws = n_ws; wd = n_wd; }                    // G:235-239
  const double thc_abs = fabs(n_thc);
  const double mws = (ws + n_ws) / 2.0;                          // G:243
  const double dws = fabs(ws - n_ws);                            // G:244
  const double dwd = fabs(modulo64((wd - n_wd) + 180.0, 360.0) - 180.0);   // G:246-247
  double sb = 0.0;
  uint8_t fl = 1;
  if (dwd < a.thresh_winddir && dws < a.thresh_windch && mws < a.thresh_wind && thc_abs > a.thresh_thc) {   // G:249-252
    const double scale_wind = (a.thresh_wind - mws) / fmax(1.0, mws);     // G:254
    const double scale_thc = (thc_abs - a.thresh_thc) / n_thc;            // G:255
    sb = scale_thc * scale_wind;
    fl = 3;
  }
  a.sb[c] = sb;
  a.ws[c] = n_ws;                                                // G:261
  a.thc[c] = n_thc;                                              // G:262
  if (a.six_hourly || a.first_step) a.wd[c] = n_wd;              // G:237, G:264-266
  if (a.flags) a.flags[c] = fl;
}

// centred crop between double planes of the same interior with different halo widths
__global__ void k64_crop_halo(double* __restrict__ dst, const double* __restrict__ src, int nx, int ny, int hd, int hs)
{
  const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
  if (x >= nx + 2 * hd) return;
  dst[(size_t)y * (nx + 2 * hd) + x] = src[(size_t)(y + hs - hd) * (nx + 2 * hs) + (x + hs - hd)];
}
