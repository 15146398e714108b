// sbd_stepq.cuh -- the whole generic timestep as ONE kernel for grids whose search radii stay small
// (N512-class global grids: radii <= 8 cells).  Included by sbd_cuda.cu inside namespace sbdk.
//
// Round 1 ran the step as two kernels that overlap through programmatic dependent launch: the window kernel
// (k_trigger, 26-28 us alone for ~1.7e7 window adds: staging, bucketing and three block barriers are 60 % of its
// instructions) and the level scan (k_scan_quad_regs, 35.9 us, HBM-bound); together 47 us, i.e. 11 us of the window
// kernel stayed exposed.  scripts/mb_scan.cu (profiles/r2_mb_scan.txt) shows the scan itself cannot get faster on
// this access pattern: a gather of the coastal quads' pressure columns costs the same ~35 us whether it touches
// 116 MB (quads), 172 MB (64-byte granules) or all 220 MB -- the cost follows the number of 128-byte lines touched,
// and runs of ~16 coastal cells touch 97 % of them.  So the step's floor is the scan, and the only thing left to
// remove is everything else.
//
// Here a scan thread -- which already owns an aligned quad of four horizontally adjacent cells -- also evaluates the
// window sums of its four cells, between issuing its first pressure loads and consuming them:
//   * the quad's window rows are read straight from the L2-resident 2-D planes (theta, gzsm, code) into
//     registers: 4 + 2 nw values per row (nw = the quad's largest radius, dispatched at compile time), shared by
//     the four cells; each cell adds its own columns in the reference's raster order with two predicated
//     scalar adds per element (T_l / T_s of generic/sea_breeze_diag.f90:200-205; the other class receives -0.0,
//     the additive identity of round-to-nearest, so the bits are the reference's);
//   * n_thc (G:216) stays in registers and meets the scan's wind speed / direction in the same thread: no
//     scratch plane, no dependent-launch handshake, no second kernel;
//   * cells outside the coastal zone only need sb_con = missing (G:176): coastal quads do it for their own
//     inactive cells, the quads without any coastal cell are a second static list handled by extra CTAs at the
//     end of the grid.
// The ~1.7e7 window adds are < 3 us of issue slots spread over a kernel that waits on HBM for 35 us.
// Needs nx % 4 == 0 (quads do not straddle rows), 16-byte aligned p,u,v, radii <= 8, minloc level search.

// window rows of one quad, radius class NW (compile-time): T_l, T_s of its four cells
template <int NW>
__device__ __forceinline__ void quad_window(const TrigArgs& a, int x, int y, const int (&nn)[4], float (&T_l)[4],
                                            float (&T_s)[4])
{
  constexpr int LW = 4 + 2 * NW;
  const int h = a.g.h;
  const int col0 = x - NW + h;                       // plane column of the first value of a row
#pragma unroll 1
  for (int dy = -NW; dy <= NW; ++dy) {
    const int gy = y + dy;
    int gyl = gy;
    if ((a.g.vh & 2) && gyl < 0) gyl = 0;            // virtual halo: domain-edge rows replicated
    if ((a.g.vh & 4) && gyl >= a.g.ny) gyl = a.g.ny - 1;
    const size_t rowoff = (size_t)(gyl + h) * a.g.ldx + col0;
    const float* trow = a.theta + rowoff;
    if (gy < 0 && a.theta_up) trow = a.theta_up + (size_t)(a.ny_up + gy + h) * a.g.ldx + col0;
    else if (gy >= a.g.ny && a.theta_dn) trow = a.theta_dn + (size_t)(gy - a.g.ny + h) * a.g.ldx + col0;
    const float* grow = a.gzsm + rowoff;
    const uint16_t* crow = a.code + rowoff;
    float tl[LW], ts[LW];
#pragma unroll
    for (int j = 0; j < LW; ++j) {
      const float t0 = __ldg(trow + j) - __ldg(grow + j);          // G:167
      const bool land = (__ldg(crow + j) >> 15) != 0;
      tl[j] = land ? t0 : -0.0f;
      ts[j] = land ? -0.0f : t0;
    }
    const int ady = abs(dy);
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int r = (nn[q] > 0 && ady <= nn[q]) ? nn[q] : -1;   // columns |d| <= r of this row belong to cell q's window
#pragma unroll
      for (int j = 0; j < LW; ++j) {
        const int d = j - NW - q;                      // compile-time column offset from cell q
        if (d >= -NW && d <= NW) add_pair_if(T_l[q], T_s[q], tl[j], ts[j], (uint32_t)(r >= (d < 0 ? -d : d)));
      }
    }
  }
}

template <int B, int NT, int MINB>
__global__ void __launch_bounds__(NT, MINB) k_step_quad(const TrigArgs a, const int* __restrict__ quads, int n_quads,
                                                  const int* __restrict__ quads_off, int n_off)
{
  const int nb_main = (n_quads + NT - 1) / NT;
  if ((int)blockIdx.x >= nb_main) {
    // quads without any coastal-zone cell: sb_con = missing (G:176), flags = 0
    const int i = ((int)blockIdx.x - nb_main) * NT + threadIdx.x;
    if (i < n_off) {
      const int m = quads_off[i];
      const float ms = a.prm.missing;
      reinterpret_cast<float4*>(a.sb)[m] = make_float4(ms, ms, ms, ms);
      if (a.flags) reinterpret_cast<uint32_t*>(a.flags)[m] = 0u;
    }
    return;
  }
  const int i = blockIdx.x * NT + threadIdx.x;
  const bool live = i < n_quads;
  const int m = quads[live ? i : n_quads - 1];
  const int c0 = m * 4;
  const int y = c0 / a.g.nx, x = c0 - y * a.g.nx;
  if (a.sync_local) {
    // multi-GPU row blocks: post "my theta of this step is in place" to the neighbours, and -- only in CTAs whose
    // quads read neighbour rows -- wait for theirs (see k_trigger)
    if (blockIdx.x == 0 && threadIdx.x == 0) {
      __threadfence_system();
      if (a.sync_up) st_release_sys(a.sync_up + 1, a.sync_seq);
      if (a.sync_dn) st_release_sys(a.sync_dn + 0, a.sync_seq);
    }
    const int need_up = (a.theta_up && y - a.g.h < 0) ? 1 : 0;
    const int need_dn = (a.theta_dn && y + a.g.h >= a.g.ny) ? 1 : 0;
    const int need = __syncthreads_or(need_up | (need_dn << 1));
    if (need) {
      if (threadIdx.x == 0) {
        if (need & 1) wait_peer_flag(a.sync_local, 0, a.sync_seq);
        if (need & 2) wait_peer_flag(a.sync_local, 1, a.sync_seq);
      }
      __syncthreads();
    }
  }
  if (!live) return;
  const size_t n2 = (size_t)a.g.nx * a.g.ny;
  const size_t stride4 = n2 >> 2;
  const float tp = a.prm.target_plev;
  const int nz = a.nz;
  const float4* pq = reinterpret_cast<const float4*>(a.p) + m;

  // the first B levels of the quad's four pressure columns are requested before anything else
  float4 d[B];
#pragma unroll
  for (int j = 0; j < B; ++j) {
    const int kk = min(j, nz - 1);
    asm volatile("ld.global.cs.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(d[j].x), "=f"(d[j].y), "=f"(d[j].z), "=f"(d[j].w)
                 : "l"(pq + (size_t)kk * stride4));
  }
  // static per-cell data and the previous state (G:243-247 need windspeed / winddir of the last step)
  const size_t so = (size_t)(y + a.g.h) * a.g.ldx + (x + a.g.h);
  int nn[4];
  bool land[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const uint32_t cd = a.code[so + j];
    nn[j] = (int)(cd & 0x7fffu);
    land[j] = (cd >> 15) != 0;
  }
  float4 ws4 = make_float4(0.f, 0.f, 0.f, 0.f), wd4 = ws4;
  if (!a.first_step) {
    ws4 = __ldcg(reinterpret_cast<const float4*>(a.ws) + m);
    wd4 = __ldcg(reinterpret_cast<const float4*>(a.wd) + m);
  }
  const uint4 cl4 = __ldg(reinterpret_cast<const uint4*>(a.cl) + m);

  // ---- window sums of the four cells (G:188-216) ----
  float T_l[4] = {0.f, 0.f, 0.f, 0.f}, T_s[4] = {0.f, 0.f, 0.f, 0.f};
  const int nw = max(max(nn[0], nn[1]), max(nn[2], nn[3]));
  const bool wrap_edge = (a.g.vh & 1) && (x < nw || x + 3 + nw >= a.g.nx);   // periodic wrap folded into the reads
  uint32_t cg_l[4] = {cl4.x, cl4.y, cl4.z, cl4.w};
  bool counted = false;
  if (!wrap_edge && nw <= 8) {
    switch (nw) {
      case 1: quad_window<1>(a, x, y, nn, T_l, T_s); break;
      case 2: quad_window<2>(a, x, y, nn, T_l, T_s); break;
      case 3: quad_window<3>(a, x, y, nn, T_l, T_s); break;
      case 4: quad_window<4>(a, x, y, nn, T_l, T_s); break;
      case 5: quad_window<5>(a, x, y, nn, T_l, T_s); break;
      case 6: quad_window<6>(a, x, y, nn, T_l, T_s); break;
      case 7: quad_window<7>(a, x, y, nn, T_l, T_s); break;
      default: quad_window<8>(a, x, y, nn, T_l, T_s); break;
    }
  } else {
    // quads at the east / west edge of a virtually periodic plane, or radii beyond 8: per-cell index-mapped sums
    counted = true;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      uint32_t c_l = 0, c_s = 0;
      T_l[j] = 0.0f; T_s[j] = 0.0f;
      if (nn[j] > 0) window_sums_global<SBD_VARIANT_GENERIC>(a, x + j, y, nn[j], T_l[j], T_s[j], c_l, c_s);
      cg_l[j] = c_l;
    }
  }
  (void)counted;
  float n_thc[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int wl = 2 * nn[j] + 1;
    const uint32_t c_l = cg_l[j], c_s = (uint32_t)(wl * wl) - c_l;
    // n_l, n_s are fp32 counters in the reference: n = n + 1 saturates at 2^24
    const float n_l = (float)min(c_l, 16777216u), n_s = (float)min(c_s, 16777216u);
    const float mul = land[j] ? 1.0f : -1.0f;                       // G:182-186
    n_thc[j] = mul * ((T_l[j] / n_l) - (T_s[j] / n_s));             // G:216
  }

  // ---- level search over the four pressure columns (G:223), B levels in flight ----
  float b0 = 0.0f, b1 = 0.0f, b2 = 0.0f, b3 = 0.0f;
  int l0 = 0, l1 = 0, l2 = 0, l3 = 0;
  int pf_lev = -1;
  for (int k0 = 0; k0 < nz; k0 += B) {
    if (k0 > 0) {
#pragma unroll
      for (int j = 0; j < B; ++j) {
        const int kk = min(k0 + j, nz - 1);
        asm volatile("ld.global.cs.v4.f32 {%0,%1,%2,%3}, [%4];"
                     : "=f"(d[j].x), "=f"(d[j].y), "=f"(d[j].z), "=f"(d[j].w)
                     : "l"(pq + (size_t)kk * stride4));
      }
    }
#pragma unroll
    for (int j = 0; j < B; ++j) {
      const int kk = k0 + j;
      if (kk < nz) {
        const float e0 = fabsf(d[j].x - tp), e1 = fabsf(d[j].y - tp), e2 = fabsf(d[j].z - tp), e3 = fabsf(d[j].w - tp);
        // minloc(abs(p - target_plev)): level 0 initialises, later levels need a strict < (G:223)
        if (kk == 0 || e0 < b0) { b0 = e0; l0 = kk; }
        if (kk == 0 || e1 < b1) { b1 = e1; l1 = kk; }
        if (kk == 0 || e2 < b2) { b2 = e2; l2 = kk; }
        if (kk == 0 || e3 < b3) { b3 = e3; l3 = kk; }
      }
    }
    if (l0 != pf_lev) {   // u, v at the best level so far: for a monotone column the minimum is final early
      pf_lev = l0;
      asm volatile("prefetch.global.L2 [%0];" ::"l"(reinterpret_cast<const float4*>(a.u + (size_t)l0 * n2) + m));
      asm volatile("prefetch.global.L2 [%0];" ::"l"(reinterpret_cast<const float4*>(a.v + (size_t)l0 * n2) + m));
    }
  }
  const int lv[4] = {l0, l1, l2, l3};
  float uu[4], vv[4];
  if (l0 == l1 && l1 == l2 && l2 == l3) {  // the usual case: one 16-byte load each
    const float4 u4 = __ldcs(reinterpret_cast<const float4*>(a.u + (size_t)l0 * n2) + m);
    const float4 v4 = __ldcs(reinterpret_cast<const float4*>(a.v + (size_t)l0 * n2) + m);
    uu[0] = u4.x; uu[1] = u4.y; uu[2] = u4.z; uu[3] = u4.w;
    vv[0] = v4.x; vv[1] = v4.y; vv[2] = v4.z; vv[3] = v4.w;
  } else {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      uu[j] = __ldcs(a.u + (size_t)lv[j] * n2 + c0 + j);
      vv[j] = __ldcs(a.v + (size_t)lv[j] * n2 + c0 + j);
    }
  }
  // ---- wind, thresholds, scaling, state update (G:225-266) ----
  const float wsp[4] = {ws4.x, ws4.y, ws4.z, ws4.w};
  const float wdp[4] = {wd4.x, wd4.y, wd4.z, wd4.w};
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const size_t c = (size_t)(c0 + j);
    if (nn[j] > 0) {
      float n_ws, n_wd;
      wind_from_uv(a, uu[j], vv[j], n_ws, n_wd);
      a.thc[c] = n_thc[j];                                            // G:262 (and G:236 on the first step)
      trigger_point_vals(a, c, n_thc[j], n_ws, n_wd, wsp[j], wdp[j]);
    } else {
      a.sb[c] = a.prm.missing;                                        // G:176
      if (a.flags) a.flags[c] = 0;
    }
  }
}
