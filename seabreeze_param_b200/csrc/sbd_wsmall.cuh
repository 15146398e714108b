// sbd_wsmall.cuh -- window-sum kernel for SMALL search radii (<= 8 cells: N512-class and ERA-Interim-class grids).
// Included by sbd_cuda.cu inside namespace sbdk, after sbd_win8.cuh (mkpair, add2, finish_cell).
//
// Round 1's k_trigger spent 26-28 us on the c2 grid for ~1.7e7 window adds: a 64x16 tile per CTA, cells bucketed by
// radius through shared-memory atomics, three block barriers, a serial prefix over 64 buckets -- 60 % of its
// instructions and most of its latency were bookkeeping, and its 769 CTAs were less than one wave.  Here:
//   * a CTA owns 32 x 8 cells (one warp per row); the tile + halo R (the plan's largest radius) is staged as ONE
//     float per cell, x = +t0 (land) / -t0 (sea) -- see sbd_win8.cuh for the signed-class form -- in 3.5 KB;
//   * one thread per cell, no bucketing, no list, ONE barrier: the cell's (2nn+1)^2 window is summed in the
//     reference's raster order (G:198-208) with one LDS + two FMNMX + one FADD2 per element, the loops unrolled for
//     the radius (switch over 1..8; neighbouring cells have neighbouring radii, so a warp runs 2-3 of the cases);
//   * ~3000 small CTAs, 8 per SM: the kernel is a few microseconds of latency, most of it under the level scan that
//     starts right behind it (programmatic dependent launch).
// A tile that stages a negative or NaN t0 takes the index-mapped global path (bit-exact for any input).
template <int VARIANT, bool SPLIT>
__global__ void __launch_bounds__(256) k_wsmall(const TrigArgs a)
{
  constexpr int TX = 32, TY = 8, NT = 256;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float* s_t = reinterpret_cast<float*>(smem_raw);
  __shared__ int s_lev, s_bad;
  const int R = a.R;
  const int W = TX + 2 * R, Hh = TY + 2 * R;
  const int WP = W | 1;                                   // odd pitch: rows start in different banks
  const int x0 = blockIdx.x * TX, y0 = blockIdx.y * TY;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  if (SPLIT && a.order == 1 && !a.pipe) asm volatile("griddepcontrol.launch_dependents;");

  if (VARIANT == SBD_VARIANT_GENERIC && a.sync_local) {  // multi-GPU post / wait, as in k_trigger
    if (blockIdx.x == 0 && blockIdx.y == 0 && tid == 0) {
      __threadfence_system();
      if (a.sync_up) st_release_sys(a.sync_up + 1, a.sync_seq);
      if (a.sync_dn) st_release_sys(a.sync_dn + 0, a.sync_seq);
    }
    const bool need_up = a.theta_up && (y0 - a.g.h < 0);
    const bool need_dn = a.theta_dn && (y0 + TY + a.g.h > a.g.ny);
    if (need_up || need_dn) {
      if (tid == 0) {
        if (need_up) wait_peer_flag(a.sync_local, 0, a.sync_seq);
        if (need_dn) wait_peer_flag(a.sync_local, 1, a.sync_seq);
      }
      __syncthreads();
    }
  }
  if (tid == 0) s_bad = 0;
  if (VARIANT == SBD_VARIANT_PYTHON && warp == NT / 32 - 1) {
    // one pressure level for the whole grid (P:228): first minimum of |p - target|, 32 levels per pass
    float best = __int_as_float(0x7f800000);
    int lev = 0x7fffffff;
    for (int k = lane; k < a.nz; k += 32) {
      const float d = fabsf(a.p[k] - a.prm.target_plev);
      if (d < best) { best = d; lev = k; }       // own levels in ascending order: first minimum kept
    }
    // NaN distances never win a strict '<' in the reference's minloc either (level 1 stays selected)
    for (int o = 16; o > 0; o >>= 1) {
      const float ob = __shfl_xor_sync(0xffffffffu, best, o);
      const int ol = __shfl_xor_sync(0xffffffffu, lev, o);
      if (ob < best || (ob == best && ol < lev)) { best = ob; lev = ol; }
    }
    if (lane == 0) s_lev = (lev == 0x7fffffff) ? 0 : lev;
  }

  // own cell first (the loads overlap the staging below)
  const int lx = lane, ly = warp;
  const int gx = x0 + lx, gy = y0 + ly;
  const bool inside = gx < a.g.nx && gy < a.g.ny;
  const size_t n2 = (size_t)a.g.nx * a.g.ny;
  const size_t c = (size_t)gy * a.g.nx + gx;
  int nn = 0;
  bool land_c = false;
  if (inside) {
    const int so = own_index<VARIANT>(a.g, gx, gy);
    const uint32_t cd = a.code[so];
    nn = (int)(cd & 0x7fffu);
    land_c = (cd >> 15) != 0;
    if (nn == 0) {  // outside the coastal zone (G:174-176) or not processed (P:165)
      if (VARIANT == SBD_VARIANT_GENERIC) {
        a.sb[c] = a.prm.missing;
      } else {
        a.out[c] = a.prm.missing;                   // P:173
        a.out[n2 + c] = a.theta[so] - a.gzsm[so];   // P:278
        a.out[2 * n2 + c] = a.ws_in[c];             // P:279
        a.out[3 * n2 + c] = a.wd_in[c];             // P:280
      }
      if (a.flags) a.flags[c] = 0;
    }
  }
  // a CTA without any coastal-zone cell is done (uniform decision: one barrier-free vote per warp + shared flag
  // would cost a barrier; the staging below is cheap enough to skip only per warp-uniform tiles via ballot)
  // ---- stage x = +-t0 of the tile + halo R ----
  int bad = 0;
  for (int r = warp; r < Hh; r += NT / 32) {
    const int sy = y0 - R + r;
    const float* trow; const float* grow; const uint16_t* crow;
    if (VARIANT == SBD_VARIANT_GENERIC) {
      int gyl = sy;
      if ((a.g.vh & 2) && gyl < 0) gyl = 0;
      if ((a.g.vh & 4) && gyl >= a.g.ny) gyl = a.g.ny - 1;
      const size_t rowoff = (size_t)min(max(gyl + a.g.h, 0), a.g.ldy - 1) * a.g.ldx;
      trow = a.theta + rowoff;
      if (sy < 0 && a.theta_up) trow = a.theta_up + (size_t)(a.ny_up + sy + a.g.h) * a.g.ldx;
      else if (sy >= a.g.ny && a.theta_dn) trow = a.theta_dn + (size_t)(sy - a.g.ny + a.g.h) * a.g.ldx;
      grow = a.gzsm + rowoff; crow = a.code + rowoff;
    } else {
      const size_t rowoff = (size_t)min(max(sy, 0), a.g.ny - 1) * a.g.nx;   // ki = clamp(ii) (P:201)
      trow = a.theta + rowoff; grow = a.gzsm + rowoff; crow = a.code + rowoff;
    }
    for (int cx = lane; cx < W; cx += 32) {
      int x;
      if (VARIANT == SBD_VARIANT_GENERIC) {
        int sx = x0 - R + cx;
        if (a.g.vh & 1) sx = (sx < 0) ? sx + a.g.nx : ((sx >= a.g.nx) ? sx - a.g.nx : sx);
        x = min(max(sx + a.g.h, 0), a.g.ldx - 1);
      } else {
        int m = x0 - R + cx + 1;   // kj = max(1, modulo(jj, nlons)) (P:202), 0-based: see nbr_index
        if (m < 0 || m >= a.g.nx) { m %= a.g.nx; if (m < 0) m += a.g.nx; }
        x = (m == 0) ? 0 : m - 1;
      }
      const float t0 = trow[x] - grow[x];                        // G:167
      if (!(t0 >= 0.0f)) bad = 1;                                // negative or NaN: signed-class form unusable
      s_t[r * WP + cx] = (crow[x] >> 15) ? t0 : -t0;
    }
  }
  if (bad) s_bad = 1;
  __syncthreads();
  if (nn == 0) return;

  float T_l = 0.0f, T_s = 0.0f;
  uint32_t c_l = 0, c_s = 0;
  if (nn > R || s_bad) {
    window_sums_global<VARIANT>(a, gx, gy, nn, T_l, T_s, c_l, c_s);
  } else {
    const int lcx = R + lx, lcy = R + ly;
    unsigned long long acc = 0ull;
    switch (nn) {
      case 1: acc = window_fixed_sc<1>(s_t, WP, lcx, lcy); break;
      case 2: acc = window_fixed_sc<2>(s_t, WP, lcx, lcy); break;
      case 3: acc = window_fixed_sc<3>(s_t, WP, lcx, lcy); break;
      case 4: acc = window_fixed_sc<4>(s_t, WP, lcx, lcy); break;
      case 5: acc = window_fixed_sc<5>(s_t, WP, lcx, lcy); break;
      case 6: acc = window_fixed_sc<6>(s_t, WP, lcx, lcy); break;
      case 7: acc = window_fixed_sc<7>(s_t, WP, lcx, lcy); break;
      case 8: acc = window_fixed_sc<8>(s_t, WP, lcx, lcy); break;
      default: {
        const int wl = 2 * nn + 1;
        for (int r = 0; r < wl; ++r) {
          const float* rp = s_t + (size_t)(lcy - nn + r) * WP + lcx - nn;
#pragma unroll 4
          for (int j = 0; j < wl; ++j) acc = add2(acc, mkpair(rp[j]));
        }
      }
    }
    unpack2(acc, T_l, T_s);
    c_l = a.cl[c];
    const int wl = 2 * nn + 1;
    c_s = (uint32_t)(wl * wl) - c_l;
  }
  finish_cell<VARIANT, SPLIT>(a, gx, gy, land_c, T_l, T_s, c_l, c_s, s_lev);
}
