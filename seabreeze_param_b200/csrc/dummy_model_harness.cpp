// dummy_model_harness.cpp -- host-side stand-in for generic/dummy_model.f90 (C++ because no
// Fortran compiler exists in this image).  It reproduces atmos_step() (D:27-37):
//     get_edges -> get_dist -> [halo fill] -> physics -> seabreeze_diag           (D:32-36, D:52-54)
// on column-major float arrays through the C-ABI only (include/sbd.h, plain pointers), exactly
// what the ISO_C_BINDING shim fortran/sea_breeze_diag_mod.f90 forwards.  Per A.6 of SURVEY.md the
// trigger receives cdist (not the 0/1 coast array, D:33-36) and halo'd theta/z/sigma.
//
// usage: dummy_model_harness <in.bin> <out.bin>
//   in : int32 nx, ny, nz, h, nsteps; float timestep; then float32 arrays
//        landfrac[ny*nx], icefrac[ny*nx], lon[nx], lat[ny], z[ny*nx], sigma[ny*nx],
//        and per step: theta[ny*nx], p[nz*ny*nx], u[nz*ny*nx], v[nz*ny*nx]
//   out: float32 mask[(ny+2h)*(nx+2h)] (coast), cdist[ny*nx], then per step
//        sb_con, windspeed, winddir, thc [ny*nx each]
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "sbd.h"

static void die(const char* what, int rc)
{
  fprintf(stderr, "dummy_model_harness: %s failed (%d): %s\n", what, rc, sbd_last_error());
  exit(1);
}
static void rd(FILE* f, void* p, size_t n) { if (fread(p, 1, n, f) != n) { fprintf(stderr, "short read\n"); exit(2); } }

static void pad(const std::vector<float>& a, std::vector<float>& out, int nx, int ny, int h)
{
  const int ldx = nx + 2 * h;
  out.assign((size_t)ldx * (ny + 2 * h), 0.0f);
  for (int y = 0; y < ny; ++y)
    for (int x = 0; x < nx; ++x) out[(size_t)(y + h) * ldx + x + h] = a[(size_t)y * nx + x];
}

int main(int argc, char** argv)
{
  if (argc != 3) { fprintf(stderr, "usage: %s in.bin out.bin\n", argv[0]); return 2; }
  FILE* fi = fopen(argv[1], "rb");
  FILE* fo = fopen(argv[2], "wb");
  if (!fi || !fo) { perror("open"); return 2; }
  int hdr[5]; float timestep;
  rd(fi, hdr, sizeof hdr); rd(fi, &timestep, 4);
  const int nx = hdr[0], ny = hdr[1], nz = hdr[2], h = hdr[3], nsteps = hdr[4];
  const size_t n2 = (size_t)nx * ny, n3 = n2 * nz, nh = (size_t)(nx + 2 * h) * (ny + 2 * h);
  std::vector<float> land(n2), ice(n2), lon(nx), lat(ny), z(n2), sigma(n2);
  rd(fi, land.data(), n2 * 4); rd(fi, ice.data(), n2 * 4); rd(fi, lon.data(), nx * 4); rd(fi, lat.data(), ny * 4);
  rd(fi, z.data(), n2 * 4); rd(fi, sigma.data(), n2 * 4);
  int rc;
  // atmos_step: call get_edges(mask, ice_frac, land_frac, halo_size)                        D:32
  std::vector<float> mask(nh, 0.0f), cdist(n2);
  if ((rc = sbd_get_edges_f32(mask.data(), ice.data(), land.data(), nx, ny, h, 1, nullptr))) die("get_edges", rc);
  // call get_dist(mask, land_frac, lon, lat, 180, cdist, halo_size)                         D:33
  if ((rc = sbd_get_dist_f32(mask.data(), land.data(), lon.data(), lat.data(), 180.0f, cdist.data(), nx, ny, h))) die("get_dist", rc);
  fwrite(mask.data(), 4, nh, fo); fwrite(cdist.data(), 4, n2, fo);
  // boundary swapping of the fields the trigger reads through its window (D header steps 2, 4)
  std::vector<float> dist_h, z_h, sig_h, theta_h;
  pad(cdist, dist_h, nx, ny, h); pad(z, z_h, nx, ny, h); pad(sigma, sig_h, nx, ny, h);
  if ((rc = sbd_swap_bounds_f32(dist_h.data(), nx, ny, h))) die("swap_bounds", rc);
  if ((rc = sbd_swap_bounds_f32(z_h.data(), nx, ny, h))) die("swap_bounds", rc);
  if ((rc = sbd_swap_bounds_f32(sig_h.data(), nx, ny, h))) die("swap_bounds", rc);
  std::vector<float> theta(n2), p(n3), u(n3), v(n3), ws(n2, 0.f), wd(n2, 0.f), thc(n2, 0.f), sb(n2, 0.f);
  for (int step = 1; step <= nsteps; ++step) {
    rd(fi, theta.data(), n2 * 4); rd(fi, p.data(), n3 * 4); rd(fi, u.data(), n3 * 4); rd(fi, v.data(), n3 * 4);
    pad(theta, theta_h, nx, ny, h);
    if ((rc = sbd_swap_bounds_f32(theta_h.data(), nx, ny, h))) die("swap_bounds", rc);
    // physics -> call seabreeze_diag(timestep, timestep_number, p,u,v, theta, z, sigma, mask, ...)   D:52-54
    if ((rc = sbd_seabreeze_diag_f32(timestep, step, p.data(), u.data(), v.data(), theta_h.data(), dist_h.data(),
                                     z_h.data(), sig_h.data(), ws.data(), wd.data(), thc.data(), sb.data(), nx, ny,
                                     nz, h, nullptr, nullptr))) die("seabreeze_diag", rc);
    fwrite(sb.data(), 4, n2, fo); fwrite(ws.data(), 4, n2, fo); fwrite(wd.data(), 4, n2, fo); fwrite(thc.data(), 4, n2, fo);
  }
  fclose(fi); fclose(fo);
  return 0;
}
