// sbd_um.cuh -- kernels behind the UM hook's entry points (UM/vn10.7/sea_breeze_diag.F90 = U).
// Included by sbd_cuda.cu inside namespace sbdk.  The UM variant keeps its fields in arrays with TWO halo widths
// (U:78-117): theta, z, sigma in tdims_s (small halo), mask / coast in tdims_l (large halo, halo_i x halo_j),
// p, u, v and the state in pdims / tdims (no halo).

// centred crop / embed between planes of the same interior with different halo widths:
// dst (nx + 2 hd) x (ny + 2 hd)  <-  src (nx + 2 hsx) x (ny + 2 hsy), hd <= min(hsx, hsy)
__global__ void k_crop_halo(float* __restrict__ dst, const float* __restrict__ src, int nx, int ny, int hd, int hsx,
                            int hsy)
{
  const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
  if (x >= nx + 2 * hd) return;
  dst[(size_t)y * (nx + 2 * hd) + x] = src[(size_t)(y + hsy - hd) * (nx + 2 * hsx) + (x + hsx - hd)];
}

// theta := theta - (gmma z) sigmoid(sigma) over a whole halo'd plane (U:210-211 overwrite theta by t0)
__global__ void k_um_t0_inplace(float* __restrict__ theta, const float* __restrict__ z, const float* __restrict__ sigma,
                                long long n, const float* __restrict__ stats, float gmma)
{
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n) return;
  const float nstd = -stats[1], r = stats[2];
  const float sm = 1.0f / (1.0f + sbd_expf(nstd * (sigma[t] - r)));
  theta[t] = theta[t] - ((gmma * z[t]) * sm);
}

// swap_bounds(field, row_len, rows, levels, halo_i, halo_j, fld_type, l_vector) of the UM (U:408-410), single
// domain: east-west periodic, north-south edge rows, separate widths per direction, `levels` planes
__global__ void k_um_halo(float* __restrict__ f, int nx, int ny, int hi, int hj, int levels)
{
  const int ldx = nx + 2 * hi, ldy = ny + 2 * hj;
  const long long per = 2LL * hj * ldx + 2LL * hi * ny;
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= per * levels) return;
  const int lev = (int)(t / per);
  const long long q = t - (long long)lev * per;
  int px, py;
  if (q < 2LL * hj * ldx) {
    const int r = (int)(q / ldx);
    px = (int)(q % ldx);
    py = (r < hj) ? r : ny + hj + (r - hj);
  } else {
    const long long w = q - 2LL * hj * ldx;
    py = hj + (int)(w / (2 * hi));
    const int c = (int)(w % (2 * hi));
    px = (c < hi) ? c : nx + hi + (c - hi);
  }
  int sx = (px - hi) % nx; if (sx < 0) sx += nx;
  const int sy = min(max(py - hj, 0), ny - 1);
  float* pl = f + (size_t)lev * ldx * ldy;
  pl[(size_t)py * ldx + px] = pl[(size_t)(sy + hj) * ldx + (sx + hi)];
}

// get_dist(landfrac, coast) of the UM (U:448-601) as a gather: one thread per interior target cell X walks the
// coast cells P of its (2 ki + 1) x (2 kj + 1) window (bit rows, ffs), with 2-D latitude / longitude
// (trignometric_mod's true_latitude / true_longitude) and the source's own longitude handling (U:550-562: the
// target's longitude goes through radians and back).  min over c == c(min a) (monotone); cells visited before /
// after the target in the raster sweep are kept apart so that the visit-time reset (U:576) is reproduced.
// Targets outside the interior are not computed: the scatter's writes into halo cells are overwritten by the
// swap_bounds at U:584.
__global__ void __launch_bounds__(128)
k_um_dist(const uint32_t* __restrict__ bits, int wpr, const float* __restrict__ landfrac, const float* __restrict__ lat2d,
          const float* __restrict__ lon2d, int nx, int ny, int ki, int kj, float maxdist, float* __restrict__ coast_l,
          int hi, int hj)
{
  const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
  if (x >= nx) return;
  const float pi = 3.1415926f;
  const float r2d = 180.0f / pi, d2r = pi / 180.0f;                   // U:509-512
  const size_t X = (size_t)y * nx + x;
  const float lamX = lon2d[X] * d2r, phiX = lat2d[X] * d2r;           // U:519-520
  const float tX = r2d * lamX;
  const float l2 = (tX > 180.0f) ? d2r * (tX - 360.0f) : lamX;        // U:558-562
  const float cX = sbd_cosf(phiX);
  const float INF = __int_as_float(0x7f800000);
  float amin_e = INF, amin_l = INF;
  const int j0 = max(0, y - kj), j1 = min(ny - 1, y + kj);
  const int i0 = max(0, x - ki), i1 = min(nx - 1, x + ki);
  for (int j = j0; j <= j1; ++j) {
    const uint32_t* brow = bits + (size_t)j * wpr;
    for (int wd = i0 >> 5; wd <= (i1 >> 5); ++wd) {
      uint32_t m = brow[wd];
      if (!m) continue;
      const int base = wd << 5;
      if (base < i0) m &= 0xffffffffu << (i0 - base);
      if (base + 31 > i1) m &= 0xffffffffu >> (base + 31 - i1);
      while (m) {
        const int b = __ffs(m) - 1;
        m &= m - 1;
        const int i = base + b;
        const size_t P = (size_t)j * nx + i;
        const float lonp = lon2d[P];
        const float l1 = (lonp > 180.0f) ? d2r * (lonp - 360.0f) : d2r * lonp;   // U:550-554
        const float phiP = lat2d[P] * d2r;
        const float dphi = phiP - phiX;                                           // U:557
        const float dlam = l1 - l2;
        const float sp = sbd_sinf(dphi / 2.0f), sl = sbd_sinf(dlam / 2.0f);
        const float a = sp * sp + (sbd_cosf(phiP) * (cX * (sl * sl)));            // U:564
        const bool earlier = (j < y) || (j == y && i <= x);
        if (earlier) amin_e = fminf(amin_e, a);
        else amin_l = fminf(amin_l, a);
      }
    }
  }
  float e = 12000.0f;
  bool far = true;
  if (amin_e < INF) {
    e = hav_c(amin_e);                                                            // U:565
    far = !(e < 12000.0f);
    if (far) e = 12000.0f;
  }
  if (fabsf(e) > 2.0f * maxdist) { e = 12000.0f; far = true; }                    // U:576
  float l = INF;
  if (amin_l < INF) l = hav_c(amin_l);
  const bool land = landfrac[X] > 0.0f;                                           // U:567
  float out;
  if (l < e) out = land ? l : -l;
  else out = far ? 12000.0f : (land ? e : -e);
  coast_l[(size_t)(y + hj) * (nx + 2 * hi) + (x + hi)] = out;                     // U:590-594
}

// get_edges(mask, icefrac, landfrac) of the UM (U:328-446) on a plane with halo_i x halo_j
__global__ void k_um_binarise(float* __restrict__ mask, const float* __restrict__ ice, const float* __restrict__ land,
                              int nx, int ny, int hi, int hj)
{
  const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
  if (x >= nx) return;
  const size_t c = (size_t)y * nx + x;
  const float ic = ice[c], ld = land[c];
  float m;
  if (ic <= 0.2f) m = (ld >= 0.5f) ? 1.0f : 0.0f;          // U:390-396
  else m = (ld + ic >= 0.5f) ? 1.0f : 0.0f;                 // U:397-402
  mask[(size_t)(y + hj) * (nx + 2 * hi) + (x + hi)] = m;
}
__global__ void k_um_sobel(const float* __restrict__ mask, float* __restrict__ coast, int nx, int ny, int hi, int hj)
{
  const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
  if (x >= nx) return;
  const int ldx = nx + 2 * hi;
  const float* r0 = mask + (size_t)(y + hj - 1) * ldx + (x + hi);
  const float* r1 = r0 + ldx;
  const float* r2 = r1 + ldx;
  const float px = (r0[1] - r0[-1]) + 2.0f * (r1[1] - r1[-1]) + (r2[1] - r2[-1]);   // U:417-424
  const float py = (r2[-1] - r0[-1]) + 2.0f * (r2[0] - r0[0]) + (r2[1] - r0[1]);
  coast[(size_t)y * nx + x] = (px != 0.0f || py != 0.0f) ? 1.0f : 0.0f;             // U:425-431
}
__global__ void k_um_embed(float* __restrict__ dst_l, const float* __restrict__ src, int nx, int ny, int hi, int hj)
{
  const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
  if (x >= nx) return;
  dst_l[(size_t)(y + hj) * (nx + 2 * hi) + (x + hi)] = src[(size_t)y * nx + x];
}
__global__ void k_um_extract(float* __restrict__ dst, const float* __restrict__ src_l, int nx, int ny, int hi, int hj)
{
  const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
  if (x >= nx) return;
  dst[(size_t)y * nx + x] = src_l[(size_t)(y + hj) * (nx + 2 * hi) + (x + hi)];
}
