// sbd_win8.cuh -- window-sum kernel for search radii beyond a few cells (0.1-degree and km-scale grids), round 2.
// Included by sbd_cuda.cu inside namespace sbdk, after TrigArgs / finish_cell / the peer-flag helpers.
//
// What is computed: for every coastal-zone cell the reference's LAST pass of its growing-window search
// (generic/sea_breeze_diag.f90:191-214 == python_wrapper/seabreezediag/seabreeze_diag_python.f90:191-218):
// T_l, T_s = sequential fp32 sums of t0 over the land / sea cells of the (2nn+1)^2 window in raster order.
// The order is kept per cell, so the bits are the reference's; what changes is how many cells share a load.
//
// The work is fixed by parity -- two dependent fp32 add chains of (2nn+1)^2 elements per cell -- so the bound is the
// FP32 pipe (148 SMs x 128 lanes; scripts/mb_pipes.cu, profiles/r2_mb_pipes.txt) and the design goal is issue slots:
// adds per instruction issued.  Measured on B200: add.rn.f32x2 (FADD2) issues at half the FADD instruction rate --
// the same flops -- and does NOT leave issue slots to the instructions next to it (scripts/mb_mix.cu), and ptxas
// will not predicate it in place, so the sums are scalar FADDs.  Round 1's k_win4 (8-byte (t_land,t_sea) pairs in
// shared memory, one CTA of 512 threads per SM at halo 54) had 58 % bank-conflict wavefronts.  This kernel:
//   * the tile holds ONE float per staged cell, x = +t0 for a land cell and -t0 for a sea cell ("signed class").
//     The pair that advances both running sums is rebuilt in registers with two FMNMX:
//     (t_land, t_sea) = (max(x,-0), max(-x,-0)); -0.0 is the additive identity of round-to-nearest
//     (y + -0.0 == y for every y, +0.0 included), so the slot of the other class is a bit-exact no-op.  Needs
//     t0 >= +0 (temperatures in K); a tile that stages a negative or NaN t0 takes the per-cell global path.
//     Half the shared memory: two CTAs per SM at halo 54 (four up to halo 31), 16-byte loads carry FOUR elements.
//   * a thread owns a strip of FOUR horizontally adjacent cells (eight: sbd_plan_set_strip_cells, measured slower)
//     and walks their windows in a skewed register window (strip_walk): one LDS.128 feeds 32 adds.  The radius is
//     1-Lipschitz, so the radii of a strip differ by <= 3: only the first / last iterations of a walked row and the
//     rows beyond the smallest radius are switched per cell; the middle loop is 86 instructions for 64 adds.
//   * the tile's halo is the largest radius that actually occurs in the tile, not the plan-wide maximum.
//   * work items (strips, single cells) are bucketed by radius, longest first, and handed to warps in chunks of 32
//     through a shared counter; single cells of wide radius go first, lone wide cells walk as strips; tiles are
//     launched heaviest first from a static list (k_tile_cost).
// Tile 64 x 32 (or 64 x 16) cells, 256 threads.  Measured history and the dead ends: DESIGN.md section 4.3.

__device__ __forceinline__ unsigned long long mkpair(float x)
{
  // (t_land, t_sea) from the signed-class value; fmaxf(+-0, -0) may return either zero: both are no-ops
  return pack2(fmaxf(x, -0.0f), fmaxf(-x, -0.0f));
}

// Walk the windows of a strip of CW horizontally adjacent cells (first cell at tile-local (lcx, lcy) of the staged
// plane) in one pass; acc[q] receives cell q's (T_l, T_s).  nq[q] = radius of cell q or 0 (nothing to sum);
// nw = largest, nlo = smallest non-zero radius.
//
// Skewed walk: a row is walked in STEPS s = 0 .. off + 2 nw, and at step s cell q adds the element of column
// a0 + s + q (a0 = the 16-byte aligned column at or before lcx - nw, off = the 0..3 columns in between).  Cell q's
// window is the step range [off + nw - nq, off + nw + nq]: every cell adds ITS OWN columns left to right, rows
// top to bottom -- the reference's raster order -- but the CW cells work on CW different elements of a sliding
// register window at the same time, so for equal radii there is no ramp and no predicate at all.  Four steps
// (one 16-byte load) make an iteration; only the first itL and the last few iterations of a row, and the rows
// beyond the smallest radius, are switched per cell by bit masks.  All loop bounds are warp-uniform (maxima /
// minima over the lanes), a lane's own extent lives in its masks.
template <int I> struct IC { static constexpr int value = I; };

// Predicated advance of one cell's two running sums.  ptxas will not predicate a packed FADD2 in place
// (it emits FADD2 into a temporary plus two SEL / MOV), and the alu pipe that would execute those runs at half rate
// like the fma pipe (scripts/mb_pipes.cu), so a switched pair-add done that way costs 4-5x an unswitched one.
// Two scalar adds under one predicate stay in place: 1 LOP3 (alu) + 2 @P FADD (fma), the cost of the plain form.
__device__ __forceinline__ void add_pair_if(float& tl, float& ts, float vl, float vs, uint32_t flag)
{
  asm("{\n\t.reg .pred q;\n\tsetp.ne.u32 q, %4, 0;\n\t@q add.rn.f32 %0, %0, %2;\n\t@q add.rn.f32 %1, %1, %3;\n\t}"
      : "+f"(tl), "+f"(ts) : "f"(vl), "f"(vs), "r"(flag));
}

// One step of four cells, each under its own predicate (bit `bit` of its mask).  The four predicates are set first and
// the eight adds follow: one predicate-latency bubble per step instead of one per cell (with a predicate per pair,
// ptxas reuses a single predicate register and every pair waits ~13 cycles for it: 40 % of the kernel's samples).
__device__ __forceinline__ void add_step4_if(float& l0, float& s0, float& l1, float& s1, float& l2, float& s2, float& l3, float& s3,
                                             float a0, float b0, float a1, float b1, float a2, float b2, float a3, float b3,
                                             uint32_t m0, uint32_t m1, uint32_t m2, uint32_t m3, uint32_t bit)
{
  asm("{\n\t.reg .pred pa, pb, pc, pd;\n\t.reg .b32 ta, tb, tc, td;\n\t"
      "and.b32 ta, %16, %20;\n\tand.b32 tb, %17, %20;\n\tand.b32 tc, %18, %20;\n\tand.b32 td, %19, %20;\n\t"
      "setp.ne.u32 pa, ta, 0;\n\tsetp.ne.u32 pb, tb, 0;\n\tsetp.ne.u32 pc, tc, 0;\n\tsetp.ne.u32 pd, td, 0;\n\t"
      "@pa add.rn.f32 %0, %0, %8;\n\t@pa add.rn.f32 %1, %1, %9;\n\t@pb add.rn.f32 %2, %2, %10;\n\t@pb add.rn.f32 %3, %3, %11;\n\t"
      "@pc add.rn.f32 %4, %4, %12;\n\t@pc add.rn.f32 %5, %5, %13;\n\t@pd add.rn.f32 %6, %6, %14;\n\t@pd add.rn.f32 %7, %7, %15;\n\t}"
      : "+f"(l0), "+f"(s0), "+f"(l1), "+f"(s1), "+f"(l2), "+f"(s2), "+f"(l3), "+f"(s3)
      : "f"(a0), "f"(b0), "f"(a1), "f"(b1), "f"(a2), "f"(b2), "f"(a3), "f"(b3), "r"(m0), "r"(m1), "r"(m2), "r"(m3), "r"(bit));
}

// Four consecutive steps of one cell under ONE predicate (a row that is inside / outside the cell's window).
__device__ __forceinline__ void add4_if(float& tl, float& ts, float l0, float s0, float l1, float s1, float l2, float s2,
                                        float l3, float s3, uint32_t flag)
{
  asm("{\n\t.reg .pred q;\n\tsetp.ne.u32 q, %10, 0;\n\t"
      "@q add.rn.f32 %0, %0, %2;\n\t@q add.rn.f32 %1, %1, %3;\n\t@q add.rn.f32 %0, %0, %4;\n\t@q add.rn.f32 %1, %1, %5;\n\t"
      "@q add.rn.f32 %0, %0, %6;\n\t@q add.rn.f32 %1, %1, %7;\n\t@q add.rn.f32 %0, %0, %8;\n\t@q add.rn.f32 %1, %1, %9;\n\t}"
      : "+f"(tl), "+f"(ts) : "f"(l0), "f"(s0), "f"(l1), "f"(s1), "f"(l2), "f"(s2), "f"(l3), "f"(s3), "r"(flag));
}

// bits k = 0..3 <-> steps base + k inside [lo, hi]
__device__ __forceinline__ uint32_t range4(int lo, int hi, int base)
{
  const int a = max(lo - base, 0), b = min(hi - base, 3);
  return (b >= a) ? ((0xfu >> (3 - (b - a))) << a) : 0u;
}

template <int CW>
__device__ __forceinline__ void strip_walk(const float* __restrict__ s_t, int WP, int lcx, int lcy, int nw, int nlo,
                                           const int (&nq)[CW], float (&accl)[CW], float (&accs)[CW], bool valid, uint32_t rt_zero)
{
  constexpr int NWIN = CW + 4;               // register window: CW + 4 consecutive elements
  constexpr int NROT = NWIN / 4;             // iterations after which the window's register names repeat
  constexpr unsigned FULL = 0xffffffffu;
  if (!valid) { nw = 0; nlo = 0; }           // idle lane of a partly filled warp: all masks empty
  const int c0 = lcx - nw, a0 = c0 & ~3, off = c0 - a0;
  const int S = off + 2 * nw + 1;            // steps per row
  const int sl = off + nw - nlo;             // first / last step at which every summing cell of this strip is active
  const int sr = off + nw + nlo;
  const int itR = valid ? ((sr + 1) >> 2) : (1 << 20);
  // warp-uniform loop structure
  const int nw_w = __reduce_max_sync(FULL, nw);
  const int nlo_w = __reduce_min_sync(FULL, valid ? nlo : (1 << 20));
  const int NIT_w = __reduce_max_sync(FULL, valid ? ((S + 3) >> 2) : 0);
  const int itL_w = __reduce_max_sync(FULL, valid ? ((sl + 3) >> 2) : 0);
  const int itR_w = __reduce_min_sync(FULL, itR);
  // a row is walked in three phases with warp-uniform bounds that are multiples of NROT (so every phase starts
  // with the register window un-rotated): [0, itA) switched per cell and step by the left masks, [itA, itB) every
  // summing cell adds all four steps (rows beyond the smallest radius: switched per cell, once per row),
  // [itB, NIT_w) switched by the right masks.  Keeping the three loops apart keeps the mask bookkeeping (shift
  // counters per cell) out of the middle loop, which is where a wide window spends its time: integer adds cost
  // more than an issue slot next to a saturated fma pipe (scripts/mb_mix.cu, profiles/r2_mb_mix.txt).
  int itA = ((itL_w + NROT - 1) / NROT) * NROT;
  int itB = (itR_w > itA) ? itA + ((itR_w - itA) / NROT) * NROT : itA;
  if (itB <= itA || itA > 8) { itA = 0; itB = 0; }   // no unswitched part: the right-hand (general) masks all the way
  uint32_t ML[CW], MR[CW];
#pragma unroll
  for (int q = 0; q < CW; ++q) {
    uint32_t ml = 0u, mr = 0u;
    if (valid && nq[q] > 0) {
      const int lo = off + nw - nq[q], hi = off + nw + nq[q];   // cell q's steps
      const int l1 = min(hi, 31);
      if (l1 >= lo) ml = (0xffffffffu >> (31 - (l1 - lo))) << lo;   // bit s, s < 32
      const int r0 = max(lo - 4 * itB, 0), r1 = min(hi - 4 * itB, 31);
      if (r1 >= r0) mr = (0xffffffffu >> (31 - (r1 - r0))) << r0;   // bit s - 4 itB
    }
    ML[q] = ml; MR[q] = mr;
  }
  const int lrow = valid ? lcy : nw_w;       // idle lanes read rows that exist
  const int lcol = valid ? a0 : 0;
  for (int dy = -nw_w; dy <= nw_w; ++dy) {
    const float4* rp = reinterpret_cast<const float4*>(s_t + (size_t)(lrow + dy) * WP + lcol);
    const int ady = abs(dy);
    const bool core_row = ady <= nlo_w;      // warp-uniform
    uint32_t pr[CW];                          // this row is inside cell q's window
#pragma unroll
    for (int q = 0; q < CW; ++q) {              // integer masks, opaque to the compiler: it would turn `& pr` back into
      pr[q] = (ady <= nq[q] && nq[q] > 0) ? 0xffffffffu : 0u;   // a select on a predicate held across the whole row,
      asm volatile("" : "+r"(pr[q]));           // and predicates are what the switched loops are short of
    }
    // register window of NWIN = CW + 4 consecutive elements, each as (t_land, t_sea); the 16-byte load of the NEXT
    // iteration is issued before this iteration's adds and converted after them (shared-memory latency under the adds)
    float wl[NWIN], ws[NWIN];
#define SBD_W_SET(i, x) { wl[i] = fmaxf((x), -0.0f); ws[i] = fmaxf(-(x), -0.0f); }
    {
      const float4 g0 = rp[0];
      SBD_W_SET(0, g0.x) SBD_W_SET(1, g0.y) SBD_W_SET(2, g0.z) SBD_W_SET(3, g0.w)
      const float4 g1 = rp[1];
      SBD_W_SET(4, g1.x) SBD_W_SET(5, g1.y) SBD_W_SET(6, g1.z) SBD_W_SET(7, g1.w)
      if (CW == 8) {
        const float4 g2 = rp[2];
        SBD_W_SET(8, g2.x) SBD_W_SET(9, g2.y) SBD_W_SET(10, g2.z) SBD_W_SET(11, g2.w)
      }
    }
    // logical window element i lives in slot (i + ROT) % NWIN; after the adds the four oldest elements are dead and
    // their slots take the prefetched group
#define SBD_W_ROTATE(gn) SBD_W_SET((0 + ROT) % NWIN, gn.x) SBD_W_SET((1 + ROT) % NWIN, gn.y) \
                         SBD_W_SET((2 + ROT) % NWIN, gn.z) SBD_W_SET((3 + ROT) % NWIN, gn.w)
    auto body_all = [&](int it, auto rot) {               // every summing cell, all four steps
      constexpr int ROT = decltype(rot)::value;
      const float4 gn = rp[it + CW / 4 + 1];              // elements 4 it + CW + 4 .. + 7: the next iteration's
#pragma unroll
      for (int k = 0; k < 4; ++k)
#pragma unroll
        for (int q = 0; q < CW; ++q) {
          accl[q] = __fadd_rn(accl[q], wl[(k + q + ROT) % NWIN]);   // G:200-202
          accs[q] = __fadd_rn(accs[q], ws[(k + q + ROT) % NWIN]);   // G:204-205
        }
      SBD_W_ROTATE(gn)
    };
    uint32_t prc[CW];                                     // pr[] again, defined just before the middle loop (below)
    auto body_row = [&](int it, auto rot) {               // all four steps, cells switched by row
      constexpr int ROT = decltype(rot)::value;
      const float4 gn = rp[it + CW / 4 + 1];
#pragma unroll
      for (int q = 0; q < CW; ++q)
        add4_if(accl[q], accs[q], wl[(q + ROT) % NWIN], ws[(q + ROT) % NWIN], wl[(1 + q + ROT) % NWIN], ws[(1 + q + ROT) % NWIN],
                wl[(2 + q + ROT) % NWIN], ws[(2 + q + ROT) % NWIN], wl[(3 + q + ROT) % NWIN], ws[(3 + q + ROT) % NWIN], prc[q]);
      SBD_W_ROTATE(gn)
    };
    auto body_mask = [&](int it, auto rot, auto mode) {   // switched per cell and step; mode 0 left, 1 right, 2 general
      constexpr int ROT = decltype(rot)::value;
      constexpr int MODE = decltype(mode)::value;
      const float4 gn = rp[it + CW / 4 + 1];
      uint32_t m[CW];
      if (MODE == 0) {
        const int sh = 4 * it;                            // < 4 itA <= 32
#pragma unroll
        for (int q = 0; q < CW; ++q) m[q] = (ML[q] >> sh) & pr[q];
      } else if (MODE == 1) {
        const int sh = 4 * (it - itB);                    // < 32: the caller's loop bound
#pragma unroll
        for (int q = 0; q < CW; ++q) m[q] = (MR[q] >> sh) & pr[q];
      } else {                                            // masks from the step ranges themselves
#pragma unroll
        for (int q = 0; q < CW; ++q)
          m[q] = valid ? (range4(off + nw - nq[q], off + nw + nq[q], 4 * it) & pr[q]) : 0u;
      }
#pragma unroll
      for (int k = 0; k < 4; ++k)
#pragma unroll
        for (int q = 0; q < CW; q += 4)
          add_step4_if(accl[q], accs[q], accl[q + 1], accs[q + 1], accl[q + 2], accs[q + 2], accl[q + 3], accs[q + 3],
                       wl[(k + q + ROT) % NWIN], ws[(k + q + ROT) % NWIN], wl[(k + q + 1 + ROT) % NWIN], ws[(k + q + 1 + ROT) % NWIN],
                       wl[(k + q + 2 + ROT) % NWIN], ws[(k + q + 2 + ROT) % NWIN], wl[(k + q + 3 + ROT) % NWIN], ws[(k + q + 3 + ROT) % NWIN],
                       m[q], m[q + 1], m[q + 2], m[q + 3], 1u << k);
      SBD_W_ROTATE(gn)
    };
    // right-hand phase: full rounds of NROT through the shifted masks while the shift stays below 32, then the general form
    const int itC = itB + min(NROT == 3 ? 6 : 8, ((NIT_w - itB) / NROT) * NROT);
    // the row predicates of the middle loop must not be set before the left-hand loop: ptxas would hoist them to the
    // top of the row and the switched loops, left with one free predicate, would wait for it at every pair of adds
    // (an empty asm is not enough -- ptxas sees through it, and it folds arithmetic zeros too; `rt_zero` is a zero it
    //  cannot know, a kernel argument, and AND-ing it with a running sum ties the predicates to the end of the
    //  left-hand loop)
#define SBD_PRC_DEFINE const uint32_t zdep = __float_as_uint(accl[0]) & rt_zero; \
    _Pragma("unroll") for (int q = 0; q < CW; ++q) prc[q] = pr[q] | zdep;
    int it = 0;
    if (NROT == 3) {
      for (; it < itA; it += 3) { body_mask(it, IC<0>(), IC<0>()); body_mask(it + 1, IC<4>(), IC<0>()); body_mask(it + 2, IC<8>(), IC<0>()); }
      if (core_row) for (; it < itB; it += 3) { body_all(it, IC<0>()); body_all(it + 1, IC<4>()); body_all(it + 2, IC<8>()); }
      else { SBD_PRC_DEFINE for (; it < itB; it += 3) { body_row(it, IC<0>()); body_row(it + 1, IC<4>()); body_row(it + 2, IC<8>()); } }
      for (; it < itC; it += 3) { body_mask(it, IC<0>(), IC<1>()); body_mask(it + 1, IC<4>(), IC<1>()); body_mask(it + 2, IC<8>(), IC<1>()); }
      for (; it + 3 <= NIT_w; it += 3) { body_mask(it, IC<0>(), IC<2>()); body_mask(it + 1, IC<4>(), IC<2>()); body_mask(it + 2, IC<8>(), IC<2>()); }
      if (it < NIT_w) { body_mask(it, IC<0>(), IC<2>()); if (it + 1 < NIT_w) body_mask(it + 1, IC<4>(), IC<2>()); }
    } else {
      for (; it < itA; it += 2) { body_mask(it, IC<0>(), IC<0>()); body_mask(it + 1, IC<4>(), IC<0>()); }
      if (core_row) for (; it < itB; it += 2) { body_all(it, IC<0>()); body_all(it + 1, IC<4>()); }
      else { SBD_PRC_DEFINE for (; it < itB; it += 2) { body_row(it, IC<0>()); body_row(it + 1, IC<4>()); } }
      for (; it < itC; it += 2) { body_mask(it, IC<0>(), IC<1>()); body_mask(it + 1, IC<4>(), IC<1>()); }
      for (; it + 2 <= NIT_w; it += 2) { body_mask(it, IC<0>(), IC<2>()); body_mask(it + 1, IC<4>(), IC<2>()); }
      if (it < NIT_w) body_mask(it, IC<0>(), IC<2>());
    }
#undef SBD_PRC_DEFINE
#undef SBD_W_ROTATE
#undef SBD_W_SET
  }
}

// one cell, compile-time radius, signed-class plane
template <int NN>
__device__ __forceinline__ unsigned long long window_fixed_sc(const float* __restrict__ s_t, int WP, int lcx, int lcy)
{
  constexpr int WL = 2 * NN + 1;
  unsigned long long acc = 0ull;
#pragma unroll 1
  for (int r = 0; r < WL; ++r) {
    const float* rp = s_t + (size_t)(lcy - NN + r) * WP + lcx - NN;
#pragma unroll
    for (int j = 0; j < WL; ++j) acc = add2(acc, mkpair(rp[j]));
  }
  return acc;
}

#ifdef SBD_TRACE_TILES
// diagnostic build only (-DSBD_TRACE_TILES, scripts/probe_tiles.py): per CTA {launch entry, SM, start ns, end ns}
__device__ unsigned long long g_tile_trace[4 * 16384];
__device__ __forceinline__ unsigned long long trace_now() { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }
#endif

template <int VARIANT, bool SPLIT, int NT, bool USE8, int TYV, int MINB = 2>
__global__ void __launch_bounds__(NT, MINB) k_win8(const TrigArgs a)
{
#ifdef SBD_TRACE_TILES
  const unsigned long long tr_t0 = trace_now();
#endif
  // TYV = 32 rows per tile; 16 (sbd_plan_set_tile_rows) halves the longest CTA of a launch, but a 64 x 16 tile has 8
  // chunks of strips for its 8 warps -- nothing to balance a long chunk against -- and stages 1.7x as much per cell:
  // measured slower everywhere (profiles/r2_probe_tiles.txt), kept as a knob
  constexpr int TX = 64, TY = TYV, NW = NT / 32;
  constexpr int N8 = TX * TY / 8, N4 = TX * TY / 4, NC = TX * TY;
  constexpr int NB = 64;        // radius buckets (radii <= 56 inside the tile)
  constexpr int MIN4 = 6;
  constexpr int LONE4 = 24;     // a lone coastal-zone cell of at least this radius is walked as a strip
  constexpr int CFIRST = 32;    // single cells of at least this radius (and the global-path ones) are handed out before the strips
  const int MIN8 = !USE8 ? (1 << 20) : (a.strip_min > 0 ? a.strip_min : 12);
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int Rmax = a.R;         // the launch's shared-memory size covers a halo of Rmax
  const int WPmax = ((TX + 2 * Rmax + 3) & ~3) + 4;   // + 4: consecutive rows start one 16-byte bank group apart
  float* s_t = reinterpret_cast<float*>(smem_raw);
  uint8_t* s_nn = reinterpret_cast<uint8_t*>(s_t + (size_t)WPmax * (TY + 2 * Rmax));   // land<<7 | wide<<6 | nn
  uint16_t* s_i8 = reinterpret_cast<uint16_t*>(s_nn + NC);    // 8-strip ids, longest first
  uint16_t* s_i4 = s_i8 + N8;                                 // 4-strip ids
  uint16_t* s_ic = s_i4 + N4;                                 // cell ids
  __shared__ int s_c8[NB], s_c4[NB], s_cc[NB], s_o8[NB], s_o4[NB], s_oc[NB];
  __shared__ int s_lev, s_n8, s_n4, s_nc, s_ncw, s_rt, s_bad, s_next;

  // launch entry: tile id | part << 20 | nparts << 26.  Heaviest tiles first; a tile whose window work exceeds a fair
  // share of an SM's time is given to `nparts` CTAs, each of which stages the tile and takes every nparts-th chunk
  const uint32_t ent = a.tile_order ? (uint32_t)a.tile_order[blockIdx.x] : ((uint32_t)blockIdx.x | (1u << 26));
  const int tile = (int)(ent & 0xfffffu), part = (int)((ent >> 20) & 0x3fu), nparts = (int)(ent >> 26);
  const int x0 = (tile % a.ntx) * TX, y0 = (tile / a.ntx) * TY;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  if (SPLIT && a.order == 1) asm volatile("griddepcontrol.launch_dependents;");

  if (VARIANT == SBD_VARIANT_GENERIC && a.sync_local && !a.xplane) {  // multi-GPU post / wait, as in k_trigger (with an x-plane: done by k_xplane)
    if (blockIdx.x == 0 && tid == 0) {
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

  if (tid < NB) { s_c8[tid] = 0; s_c4[tid] = 0; s_cc[tid] = 0; }
  if (tid == 0) { s_rt = 0; s_bad = 0; s_next = 0; }
  if (VARIANT == SBD_VARIANT_PYTHON && tid == NT - 1) {  // one level for the whole grid (P:228)
    int lev = 0;
    float best = fabsf(a.p[0] - a.prm.target_plev);
    for (int k = 1; k < a.nz; ++k) {
      const float d = fabsf(a.p[k] - a.prm.target_plev);
      if (d < best) { best = d; lev = k; }
    }
    s_lev = lev;
  }
  __syncthreads();

  // ---- phase A: own cells, row-major: radius byte; cells outside the zone are finished here ----
  const size_t n2 = (size_t)a.g.nx * a.g.ny;
  int rt = 0;
#pragma unroll
  for (int q = 0; q < NC / NT; ++q) {
    const int lid = q * NT + tid;
    const int lx = lid % TX, ly = lid / TX;
    const int gx = x0 + lx, gy = y0 + ly;
    uint8_t b = 0;
    if (gx < a.g.nx && gy < a.g.ny) {
      const int so = own_index<VARIANT>(a.g, gx, gy);
      const uint32_t cd = a.code[so];
      const int nn = (int)(cd & 0x7fffu);
      const size_t c = (size_t)gy * a.g.nx + gx;
      if (nn == 0 && part == 0 && !(VARIANT == SBD_VARIANT_PYTHON && a.prefilled)) {  // outside the coastal zone (G:174-176) or not processed (P:165)
        if (VARIANT == SBD_VARIANT_GENERIC) {
          a.sb[c] = a.prm.missing;
        } else {
          a.out[c] = a.prm.missing;                   // P:173
          a.out[n2 + c] = a.theta[so] - a.gzsm[so];   // P:278
          a.out[2 * n2 + c] = a.ws_in[c];             // P:279
          a.out[3 * n2 + c] = a.wd_in[c];             // P:280
        }
        if (a.flags) a.flags[c] = 0;
      } else if (nn != 0) {
        const bool wide = nn > Rmax || nn > 56;
        b = (uint8_t)(((cd >> 15) << 7) | (wide ? 0x40u : (uint32_t)nn));
        if (!wide) rt = max(rt, nn);
      }
    }
    s_nn[lid] = b;
  }
  rt = __reduce_max_sync(0xffffffffu, rt);
  if (lane == 0 && rt > 0) atomicMax(&s_rt, rt);
  __syncthreads();
  const int R = s_rt;                       // halo this tile needs
  const int W = TX + 2 * R, Hh = TY + 2 * R;
  const int WP = ((W + 3) & ~3) + 4;        // multiple of 4 floats (16-byte rows), + 4: rows shift by one bank group

  // ---- phase B: stage x = +-t0 of the tile + halo R, one warp per row ----
  if (R > 0 && a.xplane) {   // from the step's x-plane (k_xplane): one coalesced load per cell
    int bad = 0;
    for (int ly = warp; ly < Hh; ly += NW) {
      const float* xrow = a.xplane + (size_t)(y0 - R + ly + a.xp_R) * a.xp_ld + (x0 - R + a.xp_R);
      float* srow = s_t + (size_t)ly * WP;
#pragma unroll 6
      for (int lx = lane; lx < W; lx += 32) {
        const float xv = xrow[lx];
        if (xv != xv) bad = 1;                                     // t0 negative or NaN somewhere under this tile
        srow[lx] = xv;
      }
    }
    if (bad) s_bad = 1;
  } else if (R > 0) {
    int bad = 0;
    for (int ly = warp; ly < Hh; ly += NW) {
      const int gy = y0 - R + ly;
      const float* trow = nullptr; const float* grow = nullptr; const uint16_t* crow = nullptr;
      if (VARIANT == SBD_VARIANT_GENERIC) {
        int gyl = gy;
        if ((a.g.vh & 2) && gyl < 0) gyl = 0;
        if ((a.g.vh & 4) && gyl >= a.g.ny) gyl = a.g.ny - 1;
        const size_t rowoff = (size_t)min(max(gyl + a.g.h, 0), a.g.ldy - 1) * a.g.ldx;
        trow = a.theta + rowoff;
        if (gy < 0 && a.theta_up) trow = a.theta_up + (size_t)(a.ny_up + gy + a.g.h) * a.g.ldx;
        else if (gy >= a.g.ny && a.theta_dn) trow = a.theta_dn + (size_t)(gy - a.g.ny + a.g.h) * a.g.ldx;
        grow = a.gzsm + rowoff;
        crow = a.code + rowoff;
      } else {
        const size_t rowoff = (size_t)min(max(gy, 0), a.g.ny - 1) * a.g.nx;   // ki = clamp(ii) (P:201)
        trow = a.theta + rowoff; grow = a.gzsm + rowoff; crow = a.code + rowoff;
      }
      float* srow = s_t + (size_t)ly * WP;
#pragma unroll 4
      for (int lx = lane; lx < W; lx += 32) {
        int x;
        if (VARIANT == SBD_VARIANT_GENERIC) {
          int gx = x0 - R + lx;
          if (a.g.vh & 1) gx = (gx < 0) ? gx + a.g.nx : ((gx >= a.g.nx) ? gx - a.g.nx : gx);
          x = min(max(gx + a.g.h, 0), a.g.ldx - 1);
        } else {
          int m = x0 - R + lx + 1;   // kj = max(1, modulo(jj, nlons)) (P:202), 0-based: see nbr_index
          if (m < 0 || m >= a.g.nx) { m %= a.g.nx; if (m < 0) m += a.g.nx; }
          x = (m == 0) ? 0 : m - 1;
        }
        const float t0 = trow[x] - grow[x];                        // G:167
        if (!(t0 >= 0.0f)) bad = 1;                                // negative or NaN: signed-class form unusable
        srow[lx] = (crow[x] >> 15) ? t0 : -t0;
      }
    }
    if (bad) s_bad = 1;
  }

  // ---- phase C: classify strips (8-strip s8 <-> column sx = s8 / TY, row sy = s8 % TY: vertical neighbours in
  //      adjacent lanes); two passes over the same classification: count, then place ----
  auto classify = [&](int s8, int& kind8, int (&kind4)[2], int& nw8, int (&nw4)[2], uint2& nb8) {
    const int sx = s8 / TY, sy = s8 % TY;
    nb8 = *reinterpret_cast<const uint2*>(s_nn + sy * TX + sx * 8);
    kind8 = 0; kind4[0] = kind4[1] = 0; nw8 = 0; nw4[0] = nw4[1] = 0;
    int lo8 = 1 << 20, n8c = 0;
#pragma unroll
    for (int hlf = 0; hlf < 2; ++hlf) {
      const uint32_t w4 = hlf ? nb8.y : nb8.x;
      int lo4 = 1 << 20, n4c = 0;
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const uint32_t b = (w4 >> (8 * q)) & 0xffu;
        const int nn = (b & 0x40u) ? 0 : (int)(b & 0x3fu);
        if (nn > 0) { nw4[hlf] = max(nw4[hlf], nn); lo4 = min(lo4, nn); ++n4c; }
      }
      // (a lone coastal-zone cell of a wide radius -- the rim of the zone, where radii are largest -- still goes through
      //  the strip walk: as a single cell it would run the scalar loop, ~60 us for radius 54, at the very end of the
      //  tile with one warp alive)
      if ((n4c >= 2 || nw4[hlf] >= LONE4) && nw4[hlf] >= MIN4 && nw4[hlf] - lo4 <= 3) kind4[hlf] = 1;
      nw8 = max(nw8, nw4[hlf]); lo8 = min(lo8, lo4); n8c += n4c;
    }
    if (USE8 && n8c >= 2 && nw8 >= MIN8 && nw8 - lo8 <= 7) { kind8 = 1; kind4[0] = kind4[1] = 0; }
  };
  constexpr int S8PT = (N8 + NT - 1) / NT;   // 8-strips per thread (1 for 256 threads)
  int slot8[S8PT], slot4[S8PT][2], slotc[S8PT][8];
#pragma unroll
  for (int r = 0; r < S8PT; ++r) {
    const int s8 = r * NT + tid;
    slot8[r] = -1; slot4[r][0] = slot4[r][1] = -1;
#pragma unroll
    for (int q = 0; q < 8; ++q) slotc[r][q] = -1;
    if (s8 < N8 && (s8 % nparts) == part) {     // a split tile: this CTA owns every nparts-th 8-strip (and its cells)
      int kind8, kind4[2], nw8, nw4[2]; uint2 nb8;
      classify(s8, kind8, kind4, nw8, nw4, nb8);
      if (kind8) slot8[r] = atomicAdd(&s_c8[nw8], 1);
#pragma unroll
      for (int hlf = 0; hlf < 2; ++hlf) {
        if (kind4[hlf]) slot4[r][hlf] = atomicAdd(&s_c4[nw4[hlf]], 1);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const uint32_t b = ((hlf ? nb8.y : nb8.x) >> (8 * q)) & 0xffu;
          if (b != 0 && !kind8 && !kind4[hlf]) {
            const int bk = (b & 0x40u) ? NB - 1 : (int)(b & 0x3fu);   // wide cells first (global path: longest)
            slotc[r][hlf * 4 + q] = atomicAdd(&s_cc[bk], 1);
          }
        }
      }
    }
  }
  __syncthreads();
  if (warp == 0) {
    // exclusive prefix over the buckets, longest radius first, for the three lists (lanes 0, 1, 2)
    if (lane < 3) {
      int* cnt = lane == 0 ? s_c8 : (lane == 1 ? s_c4 : s_cc);
      int* off = lane == 0 ? s_o8 : (lane == 1 ? s_o4 : s_oc);
      int run = 0;
      for (int k = NB - 1; k >= 0; --k) { off[k] = run; run += cnt[k]; }
      if (lane == 0) s_n8 = run; else if (lane == 1) s_n4 = run; else { s_nc = run; s_ncw = off[CFIRST - 1]; }
    }
  }
  __syncthreads();
#pragma unroll
  for (int r = 0; r < S8PT; ++r) {
    const int s8 = r * NT + tid;
    if (s8 < N8 && (s8 % nparts) == part) {
      int kind8, kind4[2], nw8, nw4[2]; uint2 nb8;
      classify(s8, kind8, kind4, nw8, nw4, nb8);
      const int sx = s8 / TY, sy = s8 % TY;
      if (kind8) s_i8[s_o8[nw8] + slot8[r]] = (uint16_t)s8;
#pragma unroll
      for (int hlf = 0; hlf < 2; ++hlf) {
        if (kind4[hlf]) s_i4[s_o4[nw4[hlf]] + slot4[r][hlf]] = (uint16_t)(s8 * 2 + hlf);
#pragma unroll
        for (int q = 0; q < 4; ++q)
          if (slotc[r][hlf * 4 + q] >= 0) {
            const uint32_t b = ((hlf ? nb8.y : nb8.x) >> (8 * q)) & 0xffu;
            const int bk = (b & 0x40u) ? NB - 1 : (int)(b & 0x3fu);
            s_ic[s_oc[bk] + slotc[r][hlf * 4 + q]] = (uint16_t)(sy * TX + sx * 8 + hlf * 4 + q);
          }
      }
    }
  }
  __syncthreads();
  const int n8 = s_n8, n4 = s_n4, nc = s_nc, ncw = s_ncw;
  const bool bad_tile = s_bad != 0;

  // ---- phase D: warps take chunks of 32 consecutive items: 8-strips, then 4-strips, then cells ----
  // (a chunk may straddle two kinds; the strip walks use warp-wide reductions, so the whole warp enters them)
  const int n_items = n8 + n4 + nc;
  for (;;) {
    int base = 0;
    if (lane == 0) base = atomicAdd(&s_next, 32);
    base = __shfl_sync(0xffffffffu, base, 0);
    if (base >= n_items) break;
    // item order: the ncw single cells of wide radius (scalar loop: long, and a warp does one cell per lane), then the
    // strips, then the remaining cells -- so that nothing long is left for the end of the tile, when one warp works
    // and the CTA still holds its half of the SM
    const int ei = base + lane;
    const bool iscw = ei < ncw;
    const int e = ei - ncw;                    // index into [8-strips | 4-strips | remaining cells]
    const bool is8 = !iscw && e < n8, is4 = !iscw && !is8 && e < n8 + n4, isc = iscw || (!is8 && !is4 && ei < n_items);
    if (USE8 && __any_sync(0xffffffffu, is8 && !bad_tile)) {
      const bool valid = is8 && !bad_tile;
      const int sid = valid ? (int)s_i8[e] : 0;
      const int ssx = sid / TY, ssy = sid % TY;
      const uint2 b8 = *reinterpret_cast<const uint2*>(s_nn + ssy * TX + ssx * 8);
      int nq[8];
      int nw = 0, nlo = 1 << 20;
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const uint32_t b = ((q < 4 ? b8.x : b8.y) >> (8 * (q & 3))) & 0xffu;
        nq[q] = (b & 0x40u) ? 0 : (int)(b & 0x3fu);
        if (nq[q] > 0) { nw = max(nw, nq[q]); nlo = min(nlo, nq[q]); }
      }
      float accl[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f}, accs[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
      strip_walk<8>(s_t, WP, R + ssx * 8, R + ssy, nw, nlo, nq, accl, accs, valid, a.rt_zero);
      if (valid) {
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          const uint32_t b = ((q < 4 ? b8.x : b8.y) >> (8 * (q & 3))) & 0xffu;
          const int gx = x0 + ssx * 8 + q, gy = y0 + ssy;
          if (nq[q] > 0) {
            const float T_l = accl[q], T_s = accs[q];
            const uint32_t c_l = a.cl[(size_t)gy * a.g.nx + gx];
            const int wl = 2 * nq[q] + 1;
            finish_cell<VARIANT, SPLIT>(a, gx, gy, (b & 0x80u) != 0, T_l, T_s, c_l, (uint32_t)(wl * wl) - c_l, s_lev);
          } else if (b & 0x40u) {   // wider than the plan's halo cap: global path
            float T_l = 0.0f, T_s = 0.0f; uint32_t c_l = 0, c_s = 0;
            const int nn = (int)(a.code[own_index<VARIANT>(a.g, gx, gy)] & 0x7fffu);
            window_sums_global<VARIANT>(a, gx, gy, nn, T_l, T_s, c_l, c_s);
            finish_cell<VARIANT, SPLIT>(a, gx, gy, (b & 0x80u) != 0, T_l, T_s, c_l, c_s, s_lev);
          }
        }
      }
    }
    if (__any_sync(0xffffffffu, is4 && !bad_tile)) {
      const bool valid = is4 && !bad_tile;
      const int hid = valid ? (int)s_i4[e - n8] : 0;
      const int sid = hid >> 1, hlf = hid & 1;
      const int ssx = sid / TY, ssy = sid % TY;
      const uint32_t b4 = *reinterpret_cast<const uint32_t*>(s_nn + ssy * TX + ssx * 8 + hlf * 4);
      int nq[4];
      int nw = 0, nlo = 1 << 20;
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const uint32_t b = (b4 >> (8 * q)) & 0xffu;
        nq[q] = (b & 0x40u) ? 0 : (int)(b & 0x3fu);
        if (nq[q] > 0) { nw = max(nw, nq[q]); nlo = min(nlo, nq[q]); }
      }
      float accl[4] = {0.f, 0.f, 0.f, 0.f}, accs[4] = {0.f, 0.f, 0.f, 0.f};
      strip_walk<4>(s_t, WP, R + ssx * 8 + hlf * 4, R + ssy, nw, nlo, nq, accl, accs, valid, a.rt_zero);
      if (valid) {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const uint32_t b = (b4 >> (8 * q)) & 0xffu;
          const int gx = x0 + ssx * 8 + hlf * 4 + q, gy = y0 + ssy;
          if (nq[q] > 0) {
            const float T_l = accl[q], T_s = accs[q];
            const uint32_t c_l = a.cl[(size_t)gy * a.g.nx + gx];
            const int wl = 2 * nq[q] + 1;
            finish_cell<VARIANT, SPLIT>(a, gx, gy, (b & 0x80u) != 0, T_l, T_s, c_l, (uint32_t)(wl * wl) - c_l, s_lev);
          } else if (b & 0x40u) {
            float T_l = 0.0f, T_s = 0.0f; uint32_t c_l = 0, c_s = 0;
            const int nn = (int)(a.code[own_index<VARIANT>(a.g, gx, gy)] & 0x7fffu);
            window_sums_global<VARIANT>(a, gx, gy, nn, T_l, T_s, c_l, c_s);
            finish_cell<VARIANT, SPLIT>(a, gx, gy, (b & 0x80u) != 0, T_l, T_s, c_l, c_s, s_lev);
          }
        }
      }
    }
    if (isc) {
      const int lid = s_ic[iscw ? ei : e - n8 - n4 + ncw];
      const uint32_t b = s_nn[lid];
      const int ly = lid / TX, lx = lid - ly * TX;
      const int gx = x0 + lx, gy = y0 + ly;
      float T_l = 0.0f, T_s = 0.0f;
      uint32_t c_l = 0, c_s = 0;
      if ((b & 0x40u) || bad_tile) {
        const int nn = (int)(a.code[own_index<VARIANT>(a.g, gx, gy)] & 0x7fffu);
        window_sums_global<VARIANT>(a, gx, gy, nn, T_l, T_s, c_l, c_s);
      } else {
        const int nn = (int)(b & 0x3fu);
        const int lcx = R + lx, lcy = R + ly;
        unsigned long long acc = 0ull;
        switch (nn) {
          case 1: acc = window_fixed_sc<1>(s_t, WP, lcx, lcy); break;
          case 2: acc = window_fixed_sc<2>(s_t, WP, lcx, lcy); break;
          case 3: acc = window_fixed_sc<3>(s_t, WP, lcx, lcy); break;
          case 4: acc = window_fixed_sc<4>(s_t, WP, lcx, lcy); break;
          case 5: acc = window_fixed_sc<5>(s_t, WP, lcx, lcy); break;
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
        c_l = a.cl[(size_t)gy * a.g.nx + gx];
        const int wl = 2 * nn + 1;
        c_s = (uint32_t)(wl * wl) - c_l;
      }
      finish_cell<VARIANT, SPLIT>(a, gx, gy, (b & 0x80u) != 0, T_l, T_s, c_l, c_s, s_lev);
    }
    if ((is8 || is4) && bad_tile) {
      // bad tile (negative / NaN t0 staged): every cell of the strip takes the global path
      const int hid = is8 ? (int)s_i8[e] * 2 : (int)s_i4[e - n8];
      const int sid = hid >> 1, hlf = hid & 1;
      const int ssx = sid / TY, ssy = sid % TY;
      const int ncell = is8 ? 8 : 4;
      for (int q = 0; q < ncell; ++q) {
        const int lx = ssx * 8 + (is8 ? 0 : hlf * 4) + q;
        const uint32_t b = s_nn[ssy * TX + lx];
        if (b == 0) continue;
        const int gx = x0 + lx, gy = y0 + ssy;
        float T_l = 0.0f, T_s = 0.0f; uint32_t c_l = 0, c_s = 0;
        const int nn = (int)(a.code[own_index<VARIANT>(a.g, gx, gy)] & 0x7fffu);
        window_sums_global<VARIANT>(a, gx, gy, nn, T_l, T_s, c_l, c_s);
        finish_cell<VARIANT, SPLIT>(a, gx, gy, (b & 0x80u) != 0, T_l, T_s, c_l, c_s, s_lev);
      }
    }
  }
#ifdef SBD_TRACE_TILES
  __syncthreads();
  if (threadIdx.x == 0 && blockIdx.x < 16384) {
    unsigned smid; asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
    g_tile_trace[4 * blockIdx.x + 0] = ent; g_tile_trace[4 * blockIdx.x + 1] = smid;
    g_tile_trace[4 * blockIdx.x + 2] = tr_t0; g_tile_trace[4 * blockIdx.x + 3] = trace_now();
  }
#endif
}
