// ============================================================================
// post_kernels.cuh — reductions over the DFT accumulators the time-step path
// produces, evaluated on the device so that only the results cross PCIe
// (SURVEY.md §8(f)-1):
//   * near-to-far field transformation (src/Monitors/Near2Far.jl:40-96 green3d!,
//     :254-371 _compute_far_field_cpu, :103-247 the KernelAbstractions kernels)
//   * mode-overlap integrals            (src/Monitors/ModeMonitor.jl:345-515)
// Both read four tangential DFT monitors of a plane exactly like get_flux does
// (two-plane average in Complex{T}, then ComplexF64 arithmetic) and are
// deterministic: per-block partial sums, then an ordered final sum.
// ============================================================================
#pragma once
#include "step_kernels.cuh"

namespace khr {

struct cplx {
  double re, im;
};
__device__ __forceinline__ cplx cmul(cplx a, cplx b) { return {a.re * b.re - a.im * b.im, a.re * b.im + a.im * b.re}; }
__device__ __forceinline__ cplx cscale(cplx a, double s) { return {a.re * s, a.im * s}; }
__device__ __forceinline__ cplx cadd(cplx a, cplx b) { return {a.re + b.re, a.im + b.im}; }
__device__ __forceinline__ cplx csub(cplx a, cplx b) { return {a.re - b.re, a.im - b.im}; }
__device__ __forceinline__ cplx cconj(cplx a) { return {a.re, -a.im}; }

// sum of `NV` doubles per thread over the CTA (256 threads); result valid in thread 0
template <int NV>
__device__ __forceinline__ void block_sum(double (&v)[NV]) {
  __shared__ double ws[8][NV];
#pragma unroll
  for (int j = 0; j < NV; ++j)
    for (int o = 16; o > 0; o >>= 1) v[j] += __shfl_down_sync(0xffffffffu, v[j], o);
  if ((threadIdx.x & 31) == 0)
#pragma unroll
    for (int j = 0; j < NV; ++j) ws[threadIdx.x >> 5][j] = v[j];
  __syncthreads();
  if (threadIdx.x == 0)
#pragma unroll
    for (int j = 0; j < NV; ++j) {
      double t = 0;
      for (int q = 0; q < 8; ++q) t += ws[q][j];
      v[j] = t;
    }
}

// ----------------------------------------------------------------------------
// green3d! (Near2Far.jl:40-96): field of a unit point current of component c0 (1..3 electric
// Jx,Jy,Jz; 4..6 magnetic Mx,My,Mz) with complex amplitude f0 at x0, observed at x, homogeneous
// medium (eps, mu); near (1/r^3), intermediate (1/r^2) and far (1/r) terms.
//   kn = k*n, k = 2 pi f n, Z = sqrt(mu/eps) are hoisted by the caller.
// ----------------------------------------------------------------------------
__device__ __forceinline__ void green3d(double (&Er)[6], double (&Ei)[6], const double (&x)[3], double k, double kn,
                                        double Z, double eps, double mu, const double (&x0)[3], int c0, cplx f0) {
  const double rv0 = x[0] - x0[0], rv1 = x[1] - x0[1], rv2 = x[2] - x0[2];
  const double r = sqrt((rv0 * rv0 + rv1 * rv1) + rv2 * rv2);
  if (!(r >= 1e-20)) return;   // self-interaction guard
  const double rh[3] = {rv0 / r, rv1 / r, rv2 / r};
  const double kr = k * r;
  double sn, cs;
  sincos(kr + 1.5707963267948966, &sn, &cs);   // exp(i (k r + pi/2))
  const double amp = kn / (12.566370614359172 * r);   // k n / (4 pi r)
  const cplx expfac = cmul(cscale(f0, amp), cplx{cs, sn});
  // unit source direction p, p . r_hat and r_hat x p (generic dot / cross, as the reference)
  const int pc = (c0 - 1) % 3;
  const double p[3] = {pc == 0 ? 1.0 : 0.0, pc == 1 ? 1.0 : 0.0, pc == 2 ? 1.0 : 0.0};
  const double pdot = (p[0] * rh[0] + p[1] * rh[1]) + p[2] * rh[2];
  const double cr[3] = {rh[1] * p[2] - rh[2] * p[1], rh[2] * p[0] - rh[0] * p[2], rh[0] * p[1] - rh[1] * p[0]};
  const double ikr_inv = 1.0 / kr;          // 1/(i k r) = -i / (k r)
  const double ikr2_inv = -1.0 / (kr * kr); // 1/ikr2, ikr2 = -(k r)^2
  const cplx term1 = {1.0 + ikr2_inv, ikr_inv};                               // 1 - 1/ikr + 1/ikr2
  const cplx term2 = {(-1.0 - 3.0 * ikr2_inv) * pdot, (-3.0 * ikr_inv) * pdot};  // (-1 + 3/ikr - 3/ikr2) p.r_hat
  const cplx term3 = {1.0, ikr_inv};                                          // 1 - 1/ikr
  const bool electric = c0 <= 3;
  const double med = electric ? eps : mu;
  const cplx ef = {expfac.re / med, expfac.im / med};
  const cplx e3 = cmul(ef, term3);
#pragma unroll
  for (int j = 0; j < 3; ++j) {
    const cplx dir = cmul(ef, cadd(cscale(term1, p[j]), cscale(term2, rh[j])));
    if (electric) {
      Er[j] += dir.re; Ei[j] += dir.im;
      Er[3 + j] += e3.re * cr[j] / Z; Ei[3 + j] += e3.im * cr[j] / Z;
    } else {
      Er[j] -= e3.re * cr[j] * Z; Ei[j] -= e3.im * cr[j] * Z;
      Er[3 + j] += dir.re; Ei[3 + j] += dir.im;
    }
  }
}

template <class T>
struct N2FArgs {
  FluxArgs<T> s;        // the four monitors (E1, E2, H1, H2) and the common tangential extent
  double base[4][3];    // physical position of dft[1,1,1] of each monitor (Monitors.jl:398-420)
  double d1, d2;        // grid spacing along the two tangential axes
  double ns, eps, mu;   // normal sign, medium
  const double* obs;    // nobs x 3, (x,y,z) per point
  const double* freqs;  // nf
  int nobs, nchunk;
};

// One CTA = one (observation point, frequency, surface chunk): its threads stride over the
// chunk's surface cells, each adding the four equivalent currents of a cell
// (J = n x H, M = -n x E; Near2Far.jl:145-149, 188-192, 232-236).
template <class T>
__global__ void __launch_bounds__(256) near2far_kernel(const __grid_constant__ N2FArgs<T> a, double* __restrict__ partial) {
  const int io = blockIdx.x, kf = blockIdx.y, chunk = blockIdx.z;
  const double x[3] = {a.obs[3 * (size_t)io], a.obs[3 * (size_t)io + 1], a.obs[3 * (size_t)io + 2]};
  const double freq = a.freqs[kf];
  const double n = sqrt(a.eps * a.mu);
  const double k = 6.283185307179586 * freq * n;
  const double kn = k * n;
  const double Z = sqrt(a.mu / a.eps);
  // per normal axis: source order (field index 0 e1, 1 e2, 2 h1, 3 h2; current component; sign)
  //   z: Jx from H2 (+), Jy from H1 (-), Mx from E2 (-), My from E1 (+)
  //   x: Jy from H2 (+), Jz from H1 (-), My from E2 (-), Mz from E1 (+)
  //   y: Jz from H1 (+), Jx from H2 (-), Mz from E1 (-), Mx from E2 (+)
  int fld[4], cc[4];
  double sg[4];
  if (a.s.normal == 2) { fld[0] = 3; cc[0] = 1; sg[0] = 1; fld[1] = 2; cc[1] = 2; sg[1] = -1; fld[2] = 1; cc[2] = 4; sg[2] = -1; fld[3] = 0; cc[3] = 5; sg[3] = 1; }
  else if (a.s.normal == 0) { fld[0] = 3; cc[0] = 2; sg[0] = 1; fld[1] = 2; cc[1] = 3; sg[1] = -1; fld[2] = 1; cc[2] = 5; sg[2] = -1; fld[3] = 0; cc[3] = 6; sg[3] = 1; }
  else { fld[0] = 2; cc[0] = 3; sg[0] = 1; fld[1] = 3; cc[1] = 1; sg[1] = -1; fld[2] = 0; cc[2] = 6; sg[2] = -1; fld[3] = 1; cc[3] = 4; sg[3] = 1; }
  const long long ncell = (long long)a.s.n1 * a.s.n2;
  const long long per = (ncell + a.nchunk - 1) / a.nchunk;
  const long long q0 = per * chunk, q1 = min(ncell, q0 + per);
  double Er[6] = {0, 0, 0, 0, 0, 0}, Ei[6] = {0, 0, 0, 0, 0, 0};
  for (long long q = q0 + threadIdx.x; q < q1; q += blockDim.x) {
    const int i1 = (int)(q % a.s.n1), i2 = (int)(q / a.s.n1);
#pragma unroll
    for (int sidx = 0; sidx < 4; ++sidx) {
      const int m = fld[sidx];
      T re, im;
      flux_val(a.s, m, i1, i2, kf, re, im);
      double x0[3];
#pragma unroll
      for (int ax = 0; ax < 3; ++ax)
        x0[ax] = a.base[m][ax] + (ax == a.s.t1 ? i1 * a.d1 : (ax == a.s.t2 ? i2 * a.d2 : 0.0));
      const double sns = sg[sidx] * a.ns;   // f0 = (+-ns * F) * dA
      green3d(Er, Ei, x, k, kn, Z, a.eps, a.mu, x0, cc[sidx], cplx{(sns * (double)re) * a.s.dA, (sns * (double)im) * a.s.dA});
    }
  }
  double v[12];
#pragma unroll
  for (int j = 0; j < 6; ++j) { v[2 * j] = Er[j]; v[2 * j + 1] = Ei[j]; }
  block_sum<12>(v);
  if (threadIdx.x == 0) {
    double* o = partial + 12 * (((size_t)chunk * a.s.nf + kf) * a.nobs + io);
#pragma unroll
    for (int j = 0; j < 12; ++j) o[j] = v[j];
  }
}

// out: ComplexF64 (nobs, 6, nf) column-major, interleaved re/im (Near2Far.jl:254-258)
__global__ void near2far_finish_kernel(const double* __restrict__ partial, int nchunk, int nf, int nobs, double* __restrict__ out) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (long long)nobs * nf * 6) return;
  const int io = (int)(t % nobs);
  const int j = (int)((t / nobs) % 6);
  const int kf = (int)(t / ((long long)nobs * 6));
  double re = 0, im = 0;
  for (int c = 0; c < nchunk; ++c) {
    const double* p = partial + 12 * (((size_t)c * nf + kf) * nobs + io) + 2 * j;
    re += p[0]; im += p[1];
  }
  out[2 * t] = re;
  out[2 * t + 1] = im;
}

// ----------------------------------------------------------------------------
// Mode overlap (ModeMonitor.jl:462-498): with the mode profile already interpolated onto the
// DFT grid (ComplexF64, [4][nf][n2][n1] = e1, e2, h1, h2),
//   P_mode   = sum 0.5 real(me1 conj(mh2) - me2 conj(mh1)) dA
//   overlap± = sum ((Et1 conj(mh2) - Et2 conj(mh1)) ± (conj(me1) Ht2 - conj(me2) Ht1)) dA
// partial: [nf][nblocks][5] = P, re(o+), im(o+), re(o-), im(o-)
// ----------------------------------------------------------------------------
template <class T>
__global__ void __launch_bounds__(256) mode_overlap_kernel(const __grid_constant__ FluxArgs<T> a, const double* __restrict__ mode,
                                                           double* __restrict__ partial) {
  const int kf = blockIdx.y;
  const long long ncell = (long long)a.n1 * a.n2;
  double v[5] = {0, 0, 0, 0, 0};
  for (long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x; q < ncell; q += (long long)gridDim.x * blockDim.x) {
    const int i1 = (int)(q % a.n1), i2 = (int)(q / a.n1);
    cplx f[4], m[4];
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      T re, im;
      flux_val(a, c, i1, i2, kf, re, im);
      f[c] = cplx{(double)re, (double)im};
      const double* mp = mode + 2 * (((size_t)c * a.nf + kf) * (size_t)ncell + (size_t)q);
      m[c] = cplx{mp[0], mp[1]};
    }
    v[0] += 0.5 * csub(cmul(m[0], cconj(m[3])), cmul(m[1], cconj(m[2]))).re * a.dA;
    const cplx scm = csub(cmul(f[0], cconj(m[3])), cmul(f[1], cconj(m[2])));
    const cplx mcs = csub(cmul(cconj(m[0]), f[3]), cmul(cconj(m[1]), f[2]));
    const cplx op = cscale(cadd(scm, mcs), a.dA), om = cscale(csub(scm, mcs), a.dA);
    v[1] += op.re; v[2] += op.im; v[3] += om.re; v[4] += om.im;
  }
  block_sum<5>(v);
  if (threadIdx.x == 0)
#pragma unroll
    for (int j = 0; j < 5; ++j) partial[((size_t)kf * gridDim.x + blockIdx.x) * 5 + j] = v[j];
}

__global__ void mode_overlap_finish_kernel(const double* __restrict__ partial, int nblocks, int nf, double* __restrict__ out) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= nf * 5) return;
  const int kf = t / 5, j = t % 5;
  double s = 0;
  for (int q = 0; q < nblocks; ++q) s += partial[((size_t)kf * nblocks + q) * 5 + j];
  out[t] = s;
}

// ----------------------------------------------------------------------------
// Diffraction orders (DiffractionMonitor.jl:87-165 get_diffraction_efficiencies): the reference forms
// the full O(N^2) spatial DFT of the four tangential fields (fft2_manual, :185-197) and then reads
// (2 M + 1)^2 of its bins; here one CTA computes one needed bin (order (m, n), frequency) of all
// four fields directly:  X[k1, k2] = sum_j x[j1, j2] exp(-2 pi i (k1 j1 / n1 + k2 j2 / n2)),
// accumulated in ComplexF64 and rounded to Complex{T} like the reference's result array, then
//   power = real(Et1 conj(Ht2) - Et2 conj(Ht1)) / (n1 n2)^2   for propagating orders
// (k0^2 - kx^2 - ky^2 > 0), 0 otherwise.  out: [nf][2M+1][2M+1] (n fastest); prop: same shape, 1 = propagating.
// ----------------------------------------------------------------------------
template <class T>
struct DiffArgs {
  FluxArgs<T> s;
  int max_order;
  double L1, L2, kinc1, kinc2;
  const double* freqs;
};

template <class T>
__global__ void __launch_bounds__(256) diffraction_kernel(const __grid_constant__ DiffArgs<T> a, double* __restrict__ out,
                                                          int* __restrict__ prop) {
  const int nord = 2 * a.max_order + 1;
  const int m = (int)(blockIdx.x / nord) - a.max_order, n = (int)(blockIdx.x % nord) - a.max_order;
  const int kf = blockIdx.y;
  const int n1 = a.s.n1, n2 = a.s.n2;
  const int k1 = ((m % n1) + n1) % n1, k2 = ((n % n2) + n2) % n2;    // mod(m, n1), mod(n, n2)
  const size_t o = ((size_t)kf * nord + (size_t)(m + a.max_order)) * nord + (size_t)(n + a.max_order);
  const double k0 = 6.283185307179586 * a.freqs[kf];
  const double kx = a.kinc1 + 6.283185307179586 * m / a.L1, ky = a.kinc2 + 6.283185307179586 * n / a.L2;
  const double kz_sq = k0 * k0 - kx * kx - ky * ky;
  if (kz_sq <= 0) {   // evanescent order: skipped by the reference (block-uniform)
    if (threadIdx.x == 0) { out[o] = 0.0; prop[o] = 0; }
    return;
  }
  double v[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  const long long ncell = (long long)n1 * n2;
  for (long long q = threadIdx.x; q < ncell; q += blockDim.x) {
    const int j1 = (int)(q % n1), j2 = (int)(q / n1);
    const double phase = -6.283185307179586 * ((double)k1 * j1 / n1 + (double)k2 * j2 / n2);
    double sn, cs;
    sincos(phase, &sn, &cs);
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      T re, im;
      flux_val(a.s, c, j1, j2, kf, re, im);
      const cplx p = cmul(cplx{(double)re, (double)im}, cplx{cs, sn});
      v[2 * c] += p.re; v[2 * c + 1] += p.im;
    }
  }
  block_sum<8>(v);
  if (threadIdx.x == 0) {
    const double nf = 1.0 / ((double)n1 * n2);
    cplx f[4];
#pragma unroll
    for (int c = 0; c < 4; ++c) f[c] = cplx{(double)(T)v[2 * c] * nf, (double)(T)v[2 * c + 1] * nf};   // Complex{T} bin, * norm_factor
    out[o] = csub(cmul(f[0], cconj(f[3])), cmul(f[1], cconj(f[2]))).re;
    prop[o] = 1;
  }
}

}  // namespace khr
