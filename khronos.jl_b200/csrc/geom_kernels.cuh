// ============================================================================
// geom_kernels.cuh — geometry rasterisation and subpixel smoothing on the device
// (SURVEY.md §8(f)-3): the step in front of the time-step path that produces the
// per-voxel eps^-1 / mu^-1 / sigma_D / sigma_B arrays the kernels stream, written
// straight into the library's material layout (no host arrays, no H2D copies).
//
// Reference: src/Geometry.jl:150-246 (_rasterize_object_yrange!: objects painted last to
// first inside their bounding-box index ranges, `point in shape`), :450-605 (init_geometry),
// :795-972 (_apply_subpixel_smoothing!, _smooth_component_yrange!).  The shape predicates
// (`in`, `bounds`, `surfpt_nearby`, `level`, `volfrac`) live in GeometryPrimitives.jl, which
// is not vendored and not version-pinned by the reference (Project.toml has no Manifest): they
// are restated here from that package's published definitions for Sphere, Cuboid and Cylinder.
//
// Scheme: per Yee component grid (1) every object paints `atomicMin(owner[voxel], object index)`
// over its bounding-box voxels — work proportional to the sum of bounding-box volumes, like the
// reference, and order independent; (2) one pass turns owners into material values; (3) for the
// E grids an optional smoothing pass detects interface voxels from the owners of the six
// neighbours and rewrites eps^-1 from the fill fraction of the nearest shape surface.
// All geometry arithmetic is Float64 in the reference's operation order (no FMA contraction).
// ============================================================================
#pragma once
#include "step_kernels.cuh"

namespace khr {

constexpr int GEOM_NONE = 0x7f7f7f7f;   // owner value of a voxel no object covers (memset 0x7f)

struct GeomObj {
  int kind;              // 0 Sphere, 1 Cuboid, 2 Cylinder (r[0] = radius, r[1] = half height, ax[0..2] = unit axis)
  int pad_;
  double c[3];           // centre
  double r[3];           // Sphere: r[0] = radius; Cuboid: half sizes along its axes
  double ax[9];          // Cuboid: rows = unit axis vectors (GeometryPrimitives' b.p)
  double bmin[3], bmax[3];   // bounds(shape)
  double val[4][3];      // eps^-1, mu^-1, sigma_D, sigma_B per component, already rounded to T by the host
};

struct GeomGrid {
  double origin[3];      // get_component_origin of this Yee component (utils.jl:156-170)
  int n[3];              // local cells Nx, Ny, Nz_local
  int nzg;               // global Nz
  int z_off;             // global z index of local plane 1, minus 1
  int mpx;               // material row pitch
};

// coordinate of global cell i (1-based) along an axis: origin + (i - 1) * Δ with the product formed
// in T like the reference (_precompute_coords, Geometry.jl:351-363: Int * T -> T, then + Float64)
template <class T>
__device__ __forceinline__ double geom_coord(double origin, int i, T d) {
  return origin + (double)((T)(i - 1) * d);
}

// `point in shape`: Sphere sum(abs2, x - c) <= r^2; Cuboid all(abs.(p * (x - c)) .<= r)
__device__ __forceinline__ bool geom_contains(const GeomObj& o, const double (&x)[3]) {
  const double d0 = x[0] - o.c[0], d1 = x[1] - o.c[1], d2 = x[2] - o.c[2];
  if (o.kind == 0) return ((d0 * d0 + d1 * d1) + d2 * d2) <= o.r[0] * o.r[0];
  if (o.kind == 2) {
    // Cylinder: p = (x - c) . a; |p| <= h/2 and sum(abs2, (x - c) - p a) <= r^2
    const double p = (d0 * o.ax[0] + d1 * o.ax[1]) + d2 * o.ax[2];
    if (fabs(p) > o.r[1]) return false;
    const double q0 = d0 - p * o.ax[0], q1 = d1 - p * o.ax[1], q2 = d2 - p * o.ax[2];
    return ((q0 * q0 + q1 * q1) + q2 * q2) <= o.r[0] * o.r[0];
  }
  bool in = true;
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    const double p = (o.ax[3 * k] * d0 + o.ax[3 * k + 1] * d1) + o.ax[3 * k + 2] * d2;
    in = in && (fabs(p) <= o.r[k]);
  }
  return in;
}

struct PaintItem {
  int obj;
  int lo[3], n[3];       // first local cell and extent of the bounding-box index range on this grid
  long long first;       // first voxel of this item inside the box (items cut a box into chunks)
  int count;
};

// (1) every voxel of the item's chunk that lies in the shape: owner = min(owner, object index).
// owner has one ghost plane below and above the slab in z (local plane 0 .. Nz_local+1).
template <class T>
__global__ void __launch_bounds__(256) geom_paint_kernel(const GeomObj* __restrict__ objs, const PaintItem* __restrict__ items,
                                                         const __grid_constant__ GeomGrid g, T dx, T dy, T dz,
                                                         int* __restrict__ owner) {
  const PaintItem it = items[blockIdx.x];
  const GeomObj o = objs[it.obj];
  for (int q = threadIdx.x; q < it.count; q += blockDim.x) {
    const long long v = it.first + q;
    const int lx = (int)(v % it.n[0]), ly = (int)((v / it.n[0]) % it.n[1]), lz = (int)(v / ((long long)it.n[0] * it.n[1]));
    const int ix = it.lo[0] + lx, iy = it.lo[1] + ly, iz = it.lo[2] + lz;   // local cells; iz may be 0 or Nz_local+1
    const double x[3] = {geom_coord<T>(g.origin[0], ix, dx), geom_coord<T>(g.origin[1], iy, dy),
                         geom_coord<T>(g.origin[2], iz + g.z_off, dz)};
    if (geom_contains(o, x)) atomicMin(owner + ((size_t)(ix - 1) + (size_t)g.mpx * ((size_t)(iy - 1) + (size_t)g.n[1] * iz)), it.obj);
  }
}

// (2) owners -> material values (kind = 0 eps^-1 / 1 mu^-1 with default 1, 2 sigma_D / 3 sigma_B with default 0)
template <class T>
__global__ void __launch_bounds__(256) geom_fill_kernel(const GeomObj* __restrict__ objs, const int* __restrict__ owner,
                                                        const __grid_constant__ GeomGrid g, int kind, int comp, T* __restrict__ out) {
  const long long nvox = (long long)g.n[0] * g.n[1] * g.n[2];
  for (long long v = (long long)blockIdx.x * blockDim.x + threadIdx.x; v < nvox; v += (long long)gridDim.x * blockDim.x) {
    const int ix = (int)(v % g.n[0]), iy = (int)((v / g.n[0]) % g.n[1]), iz = (int)(v / ((long long)g.n[0] * g.n[1]));
    const int ow = owner[(size_t)ix + (size_t)g.mpx * ((size_t)iy + (size_t)g.n[1] * (iz + 1))];
    const T dflt = kind < 2 ? T(1) : T(0);
    out[(size_t)ix + (size_t)g.mpx * ((size_t)iy + (size_t)g.n[1] * iz)] = ow == GEOM_NONE ? dflt : (T)objs[ow].val[kind][comp];
  }
}

// surfpt_nearby (GeometryPrimitives): nearest surface point and (unnormalised) outward normal.
__device__ __forceinline__ void geom_surfpt(const GeomObj& o, const double (&x)[3], double (&sp)[3], double (&nout)[3]) {
  const double d[3] = {x[0] - o.c[0], x[1] - o.c[1], x[2] - o.c[2]};
  if (o.kind == 0) {
    // nout = x == c ? e1 : normalize(x - c); surface point c + r nout
    const double nr = sqrt((d[0] * d[0] + d[1] * d[1]) + d[2] * d[2]);
    if (nr == 0.0) { nout[0] = 1.0; nout[1] = 0.0; nout[2] = 0.0; }
    else { nout[0] = d[0] / nr; nout[1] = d[1] / nr; nout[2] = d[2] / nr; }
#pragma unroll
    for (int k = 0; k < 3; ++k) sp[k] = o.c[k] + o.r[0] * nout[k];
    return;
  }
  if (o.kind == 2) {
    // Cylinder: axial coordinate p, radial vector q.  Inside: the nearer of side wall and end cap; outside the
    // radius only: side wall; beyond a cap only: that cap; beyond both: the rim, normal from the rim point to x.
    const double p = (d[0] * o.ax[0] + d[1] * o.ax[1]) + d[2] * o.ax[2];
    const double q[3] = {d[0] - p * o.ax[0], d[1] - p * o.ax[1], d[2] - p * o.ax[2]};
    const double rho = sqrt((q[0] * q[0] + q[1] * q[1]) + q[2] * q[2]);
    const double sgp = copysign(1.0, p);
    const double dr = o.r[0] - rho, dh = o.r[1] - fabs(p);
    double qh[3] = {0.0, 0.0, 0.0};        // radial unit vector (any direction perpendicular to the axis on the axis itself)
    if (rho > 0.0) { qh[0] = q[0] / rho; qh[1] = q[1] / rho; qh[2] = q[2] / rho; }
    else {
      const int j = fabs(o.ax[0]) <= fabs(o.ax[1]) ? (fabs(o.ax[0]) <= fabs(o.ax[2]) ? 0 : 2) : (fabs(o.ax[1]) <= fabs(o.ax[2]) ? 1 : 2);
      double e[3] = {0.0, 0.0, 0.0};
      e[j] = 1.0;
      const double pe = o.ax[j];
      double nr = 0.0;
#pragma unroll
      for (int k = 0; k < 3; ++k) { qh[k] = e[k] - pe * o.ax[k]; nr += qh[k] * qh[k]; }
      nr = sqrt(nr);
#pragma unroll
      for (int k = 0; k < 3; ++k) qh[k] /= nr;
    }
    const bool side = (dr >= 0.0 && dh >= 0.0) ? (dr < dh) : (dh >= 0.0);   // inside: nearer surface; outside: radial overshoot only
    const bool cap = (dr >= 0.0 && dh >= 0.0) ? !side : (dr >= 0.0);
    if (side) {
#pragma unroll
      for (int k = 0; k < 3; ++k) { sp[k] = o.c[k] + p * o.ax[k] + o.r[0] * qh[k]; nout[k] = qh[k]; }
    } else if (cap) {
#pragma unroll
      for (int k = 0; k < 3; ++k) { sp[k] = x[k] + (o.r[1] * sgp - p) * o.ax[k]; nout[k] = sgp * o.ax[k]; }
    } else {
#pragma unroll
      for (int k = 0; k < 3; ++k) { sp[k] = o.c[k] + o.r[1] * sgp * o.ax[k] + o.r[0] * qh[k]; nout[k] = x[k] - sp[k]; }
    }
    return;
  }
  // Cuboid with orthonormal axes (rows of ax): d' = p (x - c); n_k = sign(d'_k) * axis_k
  double dp[3], ad[3], sg[3], dl[3];
  bool isout[3], onbnd[3];
  int nout_cnt = 0;
  bool all_on = true;
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    dp[k] = (o.ax[3 * k] * d[0] + o.ax[3 * k + 1] * d[1]) + o.ax[3 * k + 2] * d[2];
    ad[k] = fabs(dp[k]);
    sg[k] = copysign(1.0, dp[k]);
    onbnd[k] = fabs(o.r[k] - ad[k]) <= 1.4901161193847656e-08 * o.r[k];   // Base.rtoldefault(Float64) = sqrt(eps)
    isout[k] = (o.r[k] < ad[k]) || onbnd[k];
    dl[k] = o.r[k] - ad[k];                                                // distance to the face pair (cos = 1)
    nout_cnt += isout[k] ? 1 : 0;
    all_on = all_on && (!isout[k] || onbnd[k]);
  }
  double shift[3] = {0.0, 0.0, 0.0};   // in axis coordinates
  double nax[3] = {0.0, 0.0, 0.0};
  if (nout_cnt == 0) {
    // strictly inside: closest face (findmin returns the first minimum)
    int i = 0;
    if (dl[1] < dl[i]) i = 1;
    if (dl[2] < dl[i]) i = 2;
    shift[i] = dl[i] * sg[i];
    nax[i] = sg[i];
  } else {
    // outside or on the boundary in one or more directions: project those directions onto the box
#pragma unroll
    for (int k = 0; k < 3; ++k)
      if (isout[k]) shift[k] = dl[k] * sg[k];
    if (all_on) {
#pragma unroll
      for (int k = 0; k < 3; ++k) nax[k] = onbnd[k] ? sg[k] : 0.0;   // sum of the outward normals of the touched faces
    } else {
#pragma unroll
      for (int k = 0; k < 3; ++k) nax[k] = -shift[k];                 // from the surface point towards x
    }
  }
#pragma unroll
  for (int j = 0; j < 3; ++j) {
    sp[j] = x[j] + ((o.ax[j] * shift[0] + o.ax[3 + j] * shift[1]) + o.ax[6 + j] * shift[2]);
    nout[j] = (o.ax[j] * nax[0] + o.ax[3 + j] * nax[1]) + o.ax[6 + j] * nax[2];
  }
}

// level(x, shape) >= 0 <=> x inside or on the surface (Sphere: 1 - |x-c|/r; Cuboid: 1 - max(|d'|/r))
__device__ __forceinline__ bool geom_level_nonneg(const GeomObj& o, const double (&x)[3]) {
  const double d[3] = {x[0] - o.c[0], x[1] - o.c[1], x[2] - o.c[2]};
  if (o.kind == 0) return 1.0 - sqrt((d[0] * d[0] + d[1] * d[1]) + d[2] * d[2]) / o.r[0] >= 0.0;
  if (o.kind == 2) {
    // Cylinder: 1 - max(|p| / (h/2), |q| / r)
    const double p = (d[0] * o.ax[0] + d[1] * o.ax[1]) + d[2] * o.ax[2];
    const double q0 = d[0] - p * o.ax[0], q1 = d[1] - p * o.ax[1], q2 = d[2] - p * o.ax[2];
    return 1.0 - fmax(fabs(p) / o.r[1], sqrt((q0 * q0 + q1 * q1) + q2 * q2) / o.r[0]) >= 0.0;
  }
  double m = 0.0;
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    const double p = (o.ax[3 * k] * d[0] + o.ax[3 * k + 1] * d[1]) + o.ax[3 * k + 2] * d[2];
    m = fmax(m, fabs(p) / o.r[k]);
  }
  return 1.0 - m >= 0.0;
}

// volfrac(vxl, nout, r0): fraction of the box [lo, hi] inside the half-space {r : nout . (r - r0) <= 0}
// (the side opposite to the outward normal).  Exact volume of a box cut by a plane, by
// inclusion-exclusion over the corners; directions with a negligible normal component drop out.
__device__ __forceinline__ double cube3(double t) { return t > 0.0 ? t * t * t : 0.0; }
__device__ __forceinline__ double sq2(double t) { return t > 0.0 ? t * t : 0.0; }
__device__ __forceinline__ double geom_volfrac(const double (&lo)[3], const double (&hi)[3], const double (&n)[3], const double (&r0)[3]) {
  double a[3], d = 0.0, amax = 0.0;
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    const double corner = n[k] >= 0.0 ? lo[k] : hi[k];   // the corner that minimises n . r
    d += n[k] * (r0[k] - corner);
    a[k] = fabs(n[k]) * (hi[k] - lo[k]);
    amax = fmax(amax, a[k]);
  }
  if (amax == 0.0) return d >= 0.0 ? 1.0 : 0.0;
  double b[3];
  int m = 0;
#pragma unroll
  for (int k = 0; k < 3; ++k)
    if (a[k] > 1e-6 * amax) b[m++] = a[k];
  double tot = 0.0;
  for (int k = 0; k < m; ++k) tot += b[k];
  if (d <= 0.0) return 0.0;
  if (d >= tot) return 1.0;
  if (m == 1) return d / b[0];
  if (m == 2) return ((sq2(d) - sq2(d - b[0])) - sq2(d - b[1]) + sq2(d - b[0] - b[1])) / (2.0 * b[0] * b[1]);
  const double s1 = (cube3(d - b[0]) + cube3(d - b[1])) + cube3(d - b[2]);
  const double s2 = (cube3(d - b[0] - b[1]) + cube3(d - b[0] - b[2])) + cube3(d - b[1] - b[2]);
  return (((cube3(d) - s1) + s2) - cube3(d - b[0] - b[1] - b[2])) / (6.0 * b[0] * b[1] * b[2]);
}

// (3) subpixel smoothing of one eps^-1 component (_smooth_component_yrange!, Geometry.jl:878-972).
// mode 1: volume averaging, eps^-1 = 1 / <eps>; mode 2: anisotropic (Farjadpour et al. 2006):
// (1 - n_c^2) <eps^-1> + n_c^2 / <eps>, n_c the normal component along the field component.
template <class T>
__global__ void __launch_bounds__(256) geom_smooth_kernel(const GeomObj* __restrict__ objs, int nobj, const int* __restrict__ owner,
                                                          const __grid_constant__ GeomGrid g, T dx, T dy, T dz, int comp, int mode,
                                                          T* __restrict__ out, unsigned long long* __restrict__ n_smoothed) {
  const long long nvox = (long long)g.n[0] * g.n[1] * g.n[2];
  const long long v = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= nvox) return;
  const int ix = (int)(v % g.n[0]) + 1, iy = (int)((v / g.n[0]) % g.n[1]) + 1, iz = (int)(v / ((long long)g.n[0] * g.n[1])) + 1;
  auto eps_inv_at = [&](int jx, int jy, int jz) -> T {
    const int ow = owner[(size_t)(jx - 1) + (size_t)g.mpx * ((size_t)(jy - 1) + (size_t)g.n[1] * jz)];
    return ow == GEOM_NONE ? T(1) : (T)objs[ow].val[0][comp];
  };
  const T ec = eps_inv_at(ix, iy, iz);
  // fast interface detection over the six face neighbours, in the reference's order
  const int off[6][3] = {{1, 0, 0}, {-1, 0, 0}, {0, 1, 0}, {0, -1, 0}, {0, 0, 1}, {0, 0, -1}};
  bool is_if = false;
  T en = ec;
  const T rtol = T(1e-6);
#pragma unroll
  for (int q = 0; q < 6; ++q) {
    const int jx = ix + off[q][0], jy = iy + off[q][1], jz = iz + off[q][2];
    const int gz = jz + g.z_off;
    if (is_if || jx < 1 || jx > g.n[0] || jy < 1 || jy > g.n[1] || gz < 1 || gz > g.nzg) continue;
    const T nb = eps_inv_at(jx, jy, jz);
    const T ad = nb - ec < T(0) ? ec - nb : nb - ec;
    const T an = nb < T(0) ? -nb : nb, ac = ec < T(0) ? -ec : ec;
    if (ad > rtol * (an > ac ? an : ac)) { is_if = true; en = nb; }
  }
  if (!is_if) return;
  const double p[3] = {geom_coord<T>(g.origin[0], ix, dx), geom_coord<T>(g.origin[1], iy, dy), geom_coord<T>(g.origin[2], iz + g.z_off, dz)};
  const double eps_c = (double)(T(1) / ec), eps_n = (double)(T(1) / en);
  double best_d2 = 1.7976931348623157e308, bn[3] = {0.0, 0.0, 1.0}, bs[3] = {p[0], p[1], p[2]};
  int best = -1;
  for (int gi = 0; gi < nobj; ++gi) {
    const GeomObj& o = objs[gi];
    double bb = 0.0;   // squared distance to the bounding box: cannot beat the current best -> skip
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      if (p[k] < o.bmin[k]) bb += (o.bmin[k] - p[k]) * (o.bmin[k] - p[k]);
      else if (p[k] > o.bmax[k]) bb += (p[k] - o.bmax[k]) * (p[k] - o.bmax[k]);
    }
    if (bb >= best_d2) continue;
    double sp[3], no[3];
    geom_surfpt(o, p, sp, no);
    const double e0 = sp[0] - p[0], e1 = sp[1] - p[1], e2 = sp[2] - p[2];
    const double d2 = (e0 * e0 + e1 * e1) + e2 * e2;
    if (d2 < best_d2) {
      best_d2 = d2; best = gi;
#pragma unroll
      for (int k = 0; k < 3; ++k) { bn[k] = no[k]; bs[k] = sp[k]; }
    }
  }
  if (best < 0) return;
  const double nrm = sqrt((bn[0] * bn[0] + bn[1] * bn[1]) + bn[2] * bn[2]);
  double nh[3] = {0.0, 0.0, 1.0};
  if (nrm > 0.0) { nh[0] = bn[0] / nrm; nh[1] = bn[1] / nrm; nh[2] = bn[2] / nrm; }
  const double hd[3] = {(double)dx / 2, (double)dy / 2, (double)dz / 2};
  const double lo[3] = {p[0] - hd[0], p[1] - hd[1], p[2] - hd[2]}, hi[3] = {p[0] + hd[0], p[1] + hd[1], p[2] + hd[2]};
  const double f_in = geom_volfrac(lo, hi, nh, bs);
  double eps_shape, eps_bg;
  if (geom_level_nonneg(objs[best], p)) { eps_shape = eps_c; eps_bg = eps_n; }
  else { eps_shape = eps_n; eps_bg = eps_c; }
  const double eps_avg = f_in * eps_shape + (1 - f_in) * eps_bg;
  const double eps_inv_harm = f_in / eps_shape + (1 - f_in) / eps_bg;
  double r;
  if (mode == 2) {
    const double nc2 = nh[comp] * nh[comp];
    r = (1 - nc2) * eps_inv_harm + nc2 / eps_avg;
  } else {
    r = 1.0 / eps_avg;
  }
  out[(size_t)(ix - 1) + (size_t)g.mpx * ((size_t)(iy - 1) + (size_t)g.n[1] * (iz - 1))] = (T)r;
  atomicAdd(n_smoothed, 1ull);
}

}  // namespace khr
