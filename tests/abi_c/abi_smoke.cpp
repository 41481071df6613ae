// A Python-free client of libkhronos_b200.so: everything it knows about the library comes from
// include/khronos_b200.h.  It reads a problem description and the expected results from a binary
// file written by the test harness (tests/test_abi_c.py dumps them with the CPU oracle), registers
// the problem through the C ABI in the order a Julia `ccall` binder would, steps,
// reads the fields and the DFT accumulator back and compares.  Exit code 0 = PASS.
//
//   abi_smoke <case.bin>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "khronos_b200.h"

#define CHK(call)                                                                 \
  do {                                                                            \
    if ((call) != 0) {                                                            \
      std::fprintf(stderr, "FAIL %s: %s\n", #call, khr_last_error());             \
      return 2;                                                                   \
    }                                                                             \
  } while (0)

template <class T>
static bool rd(FILE* f, T* p, size_t n) { return std::fread(p, sizeof(T), n, f) == n; }

int main(int argc, char** argv) {
  if (argc < 2) { std::fprintf(stderr, "usage: abi_smoke case.bin\n"); return 64; }
  FILE* f = std::fopen(argv[1], "rb");
  if (!f) { std::fprintf(stderr, "cannot open %s\n", argv[1]); return 64; }
  int32_t hdr[8];                      // Nx, Ny, Nz, nsteps, source comp, monitor comp, nfreq, decimation
  double gd[4];                        // dx, dy, dz, dt
  if (!rd(f, hdr, 8) || !rd(f, gd, 4)) return 65;
  const int N[3] = {hdr[0], hdr[1], hdr[2]};
  const size_t ncell = (size_t)N[0] * N[1] * N[2];
  std::vector<float> sigma[2][3];
  for (int g = 0; g < 2; ++g)
    for (int a = 0; a < 3; ++a) { sigma[g][a].resize(2 * (size_t)N[a] + 1); if (!rd(f, sigma[g][a].data(), sigma[g][a].size())) return 65; }
  int32_t sbox[6], mbox[6];            // source start[3], dims[3]; monitor start[3], end[3]
  double tp[4];
  if (!rd(f, sbox, 6) || !rd(f, tp, 4)) return 65;
  std::vector<float> amp(2 * (size_t)sbox[3] * sbox[4] * sbox[5]);
  if (!rd(f, amp.data(), amp.size()) || !rd(f, mbox, 6)) return 65;
  std::vector<double> freqs((size_t)hdr[6]);
  if (!rd(f, freqs.data(), freqs.size())) return 65;
  std::vector<float> eps(ncell);       // one per-voxel eps^-1 array used for all three components
  if (!rd(f, eps.data(), ncell)) return 65;
  std::vector<float> want_ez(ncell), want_hx(ncell);
  const size_t mcell = (size_t)(mbox[3] - mbox[0] + 1) * (mbox[4] - mbox[1] + 1) * (mbox[5] - mbox[2] + 1) * (size_t)hdr[6];
  std::vector<float> want_dft(2 * mcell);
  if (!rd(f, want_ez.data(), ncell) || !rd(f, want_hx.data(), ncell) || !rd(f, want_dft.data(), want_dft.size())) return 65;
  std::fclose(f);

  khr_grid_desc g;
  g.dtype = KHR_F32;
  for (int a = 0; a < 3; ++a) { g.n[a] = N[a]; g.dl[a] = gd[a]; }
  g.dt = gd[3];
  g.z_start = 1; g.nz_local = N[2]; g.rank = 0; g.nranks = 1;
  khr_ctx* ctx = nullptr;
  CHK(khr_ctx_create(0, &g, &ctx));
  for (int grp = 0; grp < 2; ++grp)
    for (int a = 0; a < 3; ++a) CHK(khr_set_pml_sigma(ctx, grp, a, sigma[grp][a].data(), (int32_t)sigma[grp][a].size()));
  for (int d = 0; d < 3; ++d) CHK(khr_set_material_array(ctx, KHR_MAT_EPS_INV, d, eps.data()));
  int32_t sid = -1, mid = -1;
  CHK(khr_source_register(ctx, hdr[4], sbox, sbox + 3, amp.data(), KHR_TIME_CW, tp, &sid));
  CHK(khr_monitor_register(ctx, hdr[5], mbox, mbox + 3, hdr[6], freqs.data(), hdr[7], &mid));
  CHK(khr_finalize_plan(ctx));
  for (int i = 0; i < hdr[3]; ++i) CHK(khr_step(ctx, 1));          // one ccall per step!, as the shim does
  CHK(khr_sync(ctx));
  int64_t ts = 0;
  CHK(khr_get_timestep(ctx, &ts));
  std::vector<float> ez(ncell), hx(ncell), dft(2 * mcell);
  CHK(khr_field_read(ctx, 2, ez.data()));
  CHK(khr_field_read(ctx, 3, hx.data()));
  CHK(khr_monitor_read(ctx, mid, dft.data()));
  double flux_dummy = 0; (void)flux_dummy;
  auto rel = [](const std::vector<float>& a, const std::vector<float>& b) {
    double num = 0, den = 0;
    for (size_t i = 0; i < a.size(); ++i) { const double d = (double)a[i] - (double)b[i]; num += d * d; den += (double)b[i] * (double)b[i]; }
    return den > 0 ? std::sqrt(num / den) : std::sqrt(num);
  };
  const double e1 = rel(ez, want_ez), e2 = rel(hx, want_hx), e3 = rel(dft, want_dft);
  CHK(khr_ctx_destroy(ctx));
  const bool ok = ts == hdr[3] && e1 < 1e-5 && e2 < 1e-5 && e3 < 1e-5;
  std::printf("%s timestep %lld Ez rel-L2 %.3e Hx rel-L2 %.3e DFT rel-L2 %.3e (khr_version %d)\n", ok ? "PASS" : "FAIL", (long long)ts, e1, e2,
              e3, (int)khr_version());
  return ok ? 0 : 1;
}
