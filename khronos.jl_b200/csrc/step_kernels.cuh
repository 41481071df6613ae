// ============================================================================
// step_kernels.cuh — the Yee half-step kernels of libkhronos_b200 (sm_100a)
//
// One kernel family covers both half-steps of Khronos.jl's step!
// (reference: src/Kernels/Kernels.jl:164-296 step_H_fused!, :298-454
// step_E_fused!) for every voxel class:
//   * curl + constitutive update      (Helpers.jl:286-298, ReferenceKernels.jl:323-362)
//   * the C -> U -> T -> W PML cascade (Helpers.jl:30-33, 39-270, 323-366)
//   * material conductivity sigma_D/B (Helpers.jl:141-154, 273-279)
//   * current sources                 (Sources/Sources.jl:346-357)
//   * Drude/Lorentz ADE polarisation  (Dispersive.jl:25-117, fused into the E half-step)
//
// Design (B200: HBM-bound stencil, no tensor cores):
//   - every thread owns 4 consecutive x cells (one 128-bit load per array for
//     Float32), LX lanes span a 4*LX-cell row, 256/LX rows per CTA; the CTA
//     marches up z over its work item, carrying the z-neighbour plane in
//     registers so every E/H array is read from HBM once per half-step
//   - x neighbours come from warp shuffles (one scalar edge load per row)
//   - the flux fields B/D are *eliminated everywhere*: all stages are linear,
//     so the cascade is run on mu^-1/eps^-1-scaled quantities and the T stage
//     acts on the stored field itself (A where sigma_own == 0, W otherwise).
//     A PML voxel therefore moves only U (sigma_next != 0) and W (sigma_own != 0)
//     in addition to the 9/12 interior words: 4w/8w/12w extra per half-step for
//     1/2/3 PML axes instead of the reference's 10w/14w/18w.
//   - auxiliary arrays are stored only on the slabs where their sigma != 0
//   - all bypasses are value driven (sigma == 0), exactly like the reference's
//     (Helpers.jl:53, 337), so one contiguous array per GPU reproduces the
//     single-chunk semantics; the reference's 27-chunk plan is never needed.
// ============================================================================
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace khr {

constexpr int XO = 31;        // storage x offset: cell ix lives at sx = ix + XO (cell 1 is 128 B aligned)
constexpr int MAXSRC = 8;     // sources per field group that travel in the kernel parameters (more: device table, StepParams::src_ext)
constexpr int MAXPOLE = 4;    // ADE poles that travel in the kernel parameters (more: device table, StepParams::pole_ext)
constexpr int CTA = 256;

struct WorkItem {
  int x0, xw;   // first cell (x0 % 4 == 1) and number of cells in x
  int y0, yh;   // first row and number of rows (<= 256/LX)
  int z0, zn;   // first local plane and number of planes marched
  int flags;    // bit0: tile may contain a source of this group
                // bit1: the per-voxel constitutive arrays are constant over the tile (values in mu[])
                // bits 4-6: axes (x, y, z) whose PML range the tile lies in
  int lx_log2;  // lanes along x = 1 << lx_log2 (3, 4 or 5); rows per CTA = 256 >> lx_log2
  double mu[3]; // tile-uniform m^-1 per component (exact: a Float32 value is representable)
  int chunk;    // z chunk of the item (unit of the H <-> E dependency counters)
  int zmask;    // bit q: plane z0 + q lies in the z PML of either field group (pml_tma.cuh loads the z slabs there)
};

// compact index along a PML axis: cells 1..lo_w map to 0..lo_w-1, cells >= hi_base
// map to lo_w + (i - hi_base)
struct Slab {
  int lo_w, hi_base;
  __host__ __device__ inline int idx(int i) const { return i <= lo_w ? i - 1 : lo_w + (i - hi_base); }
};

template <class T>
struct SrcDesc {
  const T* amp;       // complex interleaved, extent (dx,dy,dz)
  const int* slot;    // per source voxel: index of the (component, cell) flux accumulator in StepParams::Tsrc
  int comp;           // 0..2 within the group
  int s[3], d[3];     // local start cell, extent
  T an_re, an_im;     // amplitude a(t) of this step
  T ao_re, ao_im;     // amplitude applied in the previous step (0 if none)
};

template <class T>
struct PoleDesc {
  const T* sigma;     // material layout
  const T* Pc[3];     // P^n
  T* Pp[3];           // P^{n-1}; receives P^{n+1} (host swaps afterwards)
  T g1i, g1, cp, cd;  // gamma1_inv, gamma1, coefficient of P^n, coefficient of sigma*E
};

template <class T>
struct StepParams {
  const T* A[3];      // curl operand (E for the H half-step, H for the E half-step)
  T* F[3];            // updated field
  long long plane;    // PX*PY
  int px;             // PX
  int n[3];           // local cells
  T dt, idl[3];
  const T* idv[3];    // non-uniform grid: inv(Δ[i]) per cell and axis, index i-1 (Helpers.jl:283-284), else null
  // constitutive factor
  T m_inv;
  const T* m_arr[3];  // material layout or null
  int mpx;            // material row pitch
  long long mplane;   // mpx*Ny
  // PML coefficient vectors per axis, index i-1: sigma, 1-sigma, 1/(1+sigma)
  const T* sg[3];
  const T* om[3];
  const T* ip[3];
  // auxiliary slabs: W[d] lives on the slab of axis d, U[d] on the slab of axis next(d)
  T* W[3];
  T* U[3];
  Slab slab[3];
  int cxp;            // x-slab row pitch
  int cy, cz;         // compact extents of the y / z slabs
  // material conductivity (absorbers): sigma arrays + C stage arrays (material layout)
  const T* sigD[3];
  T* C[3];
  // sources
  int nsrc;
  SrcDesc<T> src[MAXSRC];
  const SrcDesc<T>* src_ext;    // any number of sources (Sources.jl:330-340 loops over all of them): when nsrc > MAXSRC
                                // the descriptors of this half-step live in a device table refreshed by the host
  // ADE
  int npole;
  PoleDesc<T> pole[MAXPOLE];
  const PoleDesc<T>* pole_ext;  // any number of poles (Dispersive.jl:186-228): device table when npole > MAXPOLE (at most 32)
  T* Dst[3];          // D kept on dispersive / Kerr voxels (material layout), as the reference does
  const T* chi3;      // Kerr coefficient per voxel (material layout) or null; E group only
  T* Tsrc;            // B/D kept on source voxels (compact, one slot per (component, cell))
  // chain mode (one stream, programmatic dependent launch): per z chunk, done_*[c] counts the
  // work items of a field group that have finished since the last reset
  int dep_on, nchunk;
  unsigned long long* done_mine;
  const unsigned long long* done_other;
  const unsigned long long* cnt_other;   // items per chunk of the other group
  unsigned long long epoch_other;        // launches of the other group this launch depends on
  int* err_flag;
  const WorkItem* items;
};

template <class T>
struct alignas(sizeof(T) * 4) V4 {
  T v[4];
};
template <class T>
__device__ __forceinline__ V4<T> ld4(const T* p) {
  return *reinterpret_cast<const V4<T>*>(p);
}
template <class T>
__device__ __forceinline__ void st4(T* p, const V4<T>& a) {
  *reinterpret_cast<V4<T>*>(p) = a;
}
template <class T>
__device__ __forceinline__ V4<T> zero4() {
  V4<T> a;
  a.v[0] = a.v[1] = a.v[2] = a.v[3] = T(0);
  return a;
}

// L2 prefetch of the line(s) a later 128-bit load will touch: the CTA asks for plane
// iz+1 while it works on plane iz, so the marching loop waits for an L2 hit instead of a
// DRAM round trip (the loads themselves stay where they are; no registers are held).
#ifndef KHR_PREFETCH
#define KHR_PREFETCH 1
#endif
#ifndef KHR_PF_DIST
#define KHR_PF_DIST 1   // planes ahead
#endif
__device__ __forceinline__ void pf_l2(const void* p) {
#if KHR_PREFETCH
  asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
#endif
}

__device__ __forceinline__ unsigned long long ld_acquire_u64(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}

template <class T>
struct Co {  // PML coefficients of one cell along one axis
  T s, om, ip;
};

// ----------------------------------------------------------------------------
// One component of the general cascade for the 4 cells of a thread, on registers
// only (the caller batches every load before and every store after, so that the
// memory operations of the three components overlap instead of serialising).
//   k      : m^-1 * K (scaled curl increment)
//   cn/cp/co: coefficients along next/prev/own axis (per element)
//   useU/useW: thread-level "any sigma != 0" on the next/own axis
//   a      : field values (in: old, out: new); u, w, c: auxiliary values (in/out)
//   s_old/s_new: m^-1-scaled additive terms riding on the stored field
//                (source S and polarisation -P) of the previous / this step
//   sd     : 0.5*dt*sigma_D per element (0 when absent)
// ----------------------------------------------------------------------------
template <class T, bool EXTRAS>
__device__ __forceinline__ void cascade(const T (&k)[4], const Co<T> (&cn)[4], const Co<T> (&cp)[4],
                                        const Co<T> (&co)[4], bool useU, bool useW, V4<T>& u, V4<T>& w, V4<T>& a,
                                        const T (&s_old)[4], const T (&s_new)[4], bool has_sd, const T (&sd)[4],
                                        V4<T>& c, const bool (&valid)[4]) {
  T in[4];
  T omt[4], ipt[4];
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    in[e] = k[e];
    omt[e] = cp[e].om;
    ipt[e] = cp[e].ip;
  }
  if (EXTRAS && has_sd) {
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      if (sd[e] != T(0)) {
        if (cn[e].s != T(0) || cp[e].s != T(0)) {
          // C stage (Helpers.jl:60-68): C <- ((1-sD)C + K)/(1+sD); dC feeds the next stage
          T c_old = c.v[e];
          T c_new = ((T(1) - sd[e]) * c_old + in[e]) / (T(1) + sd[e]);
          in[e] = c_new - c_old;
          if (valid[e]) c.v[e] = c_new;
        } else {
          // single stage (Helpers.jl:141-154): T <- ((1-sD)T + K)/(1+sD)
          omt[e] = T(1) - sd[e];
          ipt[e] = T(1) / (T(1) + sd[e]);
        }
      }
    }
  }
  if (useU) {
    // U stage (Helpers.jl:55-56): U <- ((1-sn)U + in)/(1+sn); bypassed where sn == 0
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      T un = (cn[e].om * u.v[e] + in[e]) * cn[e].ip;
      bool on = cn[e].s != T(0);
      in[e] = on ? (un - u.v[e]) : in[e];
      if (on && valid[e]) u.v[e] = un;
    }
  }
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    bool won = useW && (co[e].s != T(0));
    // T stage on the stored (scaled) flux: t = m^-1 (T + S - P); the additive
    // S/P parts are not damped by the reference, so peel them off and put the
    // new ones back (Helpers.jl:332-335)
    T t_old = won ? w.v[e] : a.v[e];
    T t_new;
    if constexpr (EXTRAS) t_new = (omt[e] * (t_old - s_old[e]) + in[e]) * ipt[e] + s_new[e];
    else t_new = (omt[e] * t_old + in[e]) * ipt[e];
    // W stage (Helpers.jl:336-343)
    T a_new = won ? ((a.v[e] + (T(1) + co[e].s) * t_new) - co[e].om * t_old) : t_new;
    if (valid[e]) {
      a.v[e] = a_new;
      if (won) w.v[e] = t_new;
    }
  }
}

// ----------------------------------------------------------------------------
// The half-step kernel.
//   GROUP 0: H from curl E (idx_curl = +1); GROUP 1: E from curl H (idx_curl = -1)
//   MODE 0: pure interior; 1: + PML cascade; 2: + sigma_D/B, sources, ADE poles
//   MARR   : 0 scalar m^-1; 1 per-voxel m^-1 arrays (eps^-1 or mu^-1); 2 per-voxel arrays that are
//            constant over every tile of the launch (the planner puts such tiles in their own table):
//            the three values ride in the work item, no loads, no vector registers
// The row width (lanes along x) is a per-item power of two, 8/16/32.
// ----------------------------------------------------------------------------
template <class T, int MODE, int AXM>
constexpr int min_ctas() {
  // registers/thread budget: 64K regs per SM, 256-thread CTAs
#ifndef KHR_MINCTA_M0
#define KHR_MINCTA_M0 3
#endif
#ifndef KHR_MINCTA_M1
#define KHR_MINCTA_M1 2
#endif
#ifndef KHR_MINCTA_M1S
#define KHR_MINCTA_M1S 3   // single-axis PML tiles
#endif
#ifndef KHR_MINCTA_M2
#define KHR_MINCTA_M2 1    // sources / conductivity / ADE tiles
#endif
  constexpr bool single = (AXM == 1 || AXM == 2 || AXM == 4);
  return sizeof(T) == 4 ? (MODE == 0 ? KHR_MINCTA_M0 : (MODE == 1 ? (single ? KHR_MINCTA_M1S : KHR_MINCTA_M1) : KHR_MINCTA_M2)) : 1;
}

//   AXM    : bit mask of the axes whose sigma may be non-zero inside the work item (the planner
//            cuts the items at the PML faces).  Coefficients of the other axes are the
//            compile-time constants (0, 1, 1), so their stages fold away exactly (x*1 == x):
//            a single-axis PML tile carries one W and one U array and nothing else.
//   NU     : non-uniform grid — the curl uses inv(Δx[ix]), inv(Δy[iy]), inv(Δz[iz]) of the updated cell
//            (get_inv_dx(Δ::AbstractVector, i), Helpers.jl:283-291) instead of three scalars
template <class T, int GROUP, int MODE, int MARR, int AXM = 7, bool NU = false>
__device__ __forceinline__ void step_body(const StepParams<T>& p, const WorkItem& it_in) {
  constexpr int IC = (GROUP == 0) ? 1 : -1;
  constexpr bool GENERAL = MODE >= 1;   // PML cascade
  constexpr bool EXTRAS = MODE == 2;    // + sources, sigma_D/B, ADE poles
  constexpr bool PXA = (AXM & 1) != 0, PYA = (AXM & 2) != 0, PZA = (AXM & 4) != 0;
  // chain mode: the next kernel of the stream may start as soon as every CTA of this one is
  // resident; data dependencies are handled by the chunk counters below (no-op otherwise)
  asm volatile("griddepcontrol.launch_dependents;");
  const WorkItem it = it_in;
  const int lxl = it.lx_log2;
  const int LX = 1 << lxl;
  const int lane_x = threadIdx.x & (LX - 1);
  const int row = threadIdx.x >> lxl;
  const int gx = it.x0 + 4 * lane_x;
  const int iy = it.y0 + row;
  const int lanes = (it.xw + 3) >> 2;
  const bool act = (lane_x < lanes) && (row < it.yh);
  // the lane that cannot get its x neighbour from a shuffle
  const bool edge = (GROUP == 0) ? (lane_x == lanes - 1) : (lane_x == 0);
  bool valid[4];
#pragma unroll
  for (int e = 0; e < 4; ++e) valid[e] = (4 * lane_x + e) < it.xw;

  const int fo = (gx + XO) + p.px * iy;                    // in-plane field offset
  const int mo = (gx - 1) + p.mpx * (iy - 1);              // in-plane material offset
  const T* __restrict__ Ax = p.A[0];
  const T* __restrict__ Ay = p.A[1];
  const T* __restrict__ Az = p.A[2];
  const T dt = p.dt;
  T idx_ = p.idl[0], idy_ = p.idl[1], idz_ = p.idl[2];
  T idxv[4] = {idx_, idx_, idx_, idx_};
  if constexpr (NU) {
    if (act) {
      const V4<T> v = ld4(p.idv[0] + gx - 1);
#pragma unroll
      for (int e = 0; e < 4; ++e) idxv[e] = v.v[e];
      idy_ = p.idv[1][iy - 1];
    }
  }

  // ---- per-thread constants of the general path ----
  Co<T> cx[4], cyc;
  bool hasx = false, hasy = false;
  int xs_off = 0, ys_off = 0;  // in-slab offsets (without the z part)
  if constexpr (GENERAL) {
#pragma unroll
    for (int e = 0; e < 4; ++e) { cx[e].s = T(0); cx[e].om = T(1); cx[e].ip = T(1); }
    cyc.s = T(0); cyc.om = T(1); cyc.ip = T(1);
    if (act) {
      if constexpr (PXA) {
        V4<T> s = ld4(p.sg[0] + gx - 1), o = ld4(p.om[0] + gx - 1), i = ld4(p.ip[0] + gx - 1);
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          cx[e].s = s.v[e]; cx[e].om = o.v[e]; cx[e].ip = i.v[e];
          hasx |= (s.v[e] != T(0));
        }
        if (hasx) xs_off = p.slab[0].idx(gx) + p.cxp * (iy - 1);
      }
      if constexpr (PYA) {
        cyc.s = p.sg[1][iy - 1]; cyc.om = p.om[1][iy - 1]; cyc.ip = p.ip[1][iy - 1];
        hasy = cyc.s != T(0);
        if (hasy) ys_off = (gx - 1) + p.mpx * p.slab[1].idx(iy);
      }
    }
  }
  // sources: which table entries can touch this tile (block-uniform mask)
  unsigned srcmask = 0;       // bit min(q, 31): sources >= 31 share the last bit and are box-tested per plane
  if constexpr (EXTRAS) {
    if (it.flags & 1) {
      for (int q = 0; q < p.nsrc; ++q) {
        const SrcDesc<T>& s = p.src_ext ? p.src_ext[q] : p.src[q];
        bool hit = (s.s[0] < it.x0 + it.xw) && (s.s[0] + s.d[0] > it.x0) && (s.s[1] < it.y0 + it.yh) &&
                   (s.s[1] + s.d[1] > it.y0) && (s.s[2] < it.z0 + it.zn) && (s.s[2] + s.d[2] > it.z0);
        if (hit) srcmask |= 1u << (q < 31 ? q : 31);
      }
    }
  }

  // z coefficients are fetched one plane ahead (they gate the z-slab loads)
  Co<T> czn;
  czn.s = T(0); czn.om = T(1); czn.ip = T(1);
  if constexpr (GENERAL && PZA) { czn.s = p.sg[2][it.z0 - 1]; czn.om = p.om[2][it.z0 - 1]; czn.ip = p.ip[2][it.z0 - 1]; }
  const int z_end = it.z0 + it.zn;
  const bool m_uniform = MARR == 1 && (it.flags & 2) != 0;  // block-uniform
  const T mu0 = (T)it.mu[0], mu1 = (T)it.mu[1], mu2 = (T)it.mu[2];   // MARR == 2
  const bool yedge = (GROUP == 0) ? (row == it.yh - 1) : (row == 0);  // row whose y neighbour belongs to another CTA

  // everything plane zp will load from HBM, requested into L2 ahead of time
  auto prefetch_plane = [&](int zp) {
    const long long nb = p.plane * (long long)zp + fo;
    pf_l2(Ax + nb + (GROUP == 0 ? p.plane : 0));
    pf_l2(Ay + nb + (GROUP == 0 ? p.plane : 0));
    pf_l2(Az + nb);
    if (yedge) { pf_l2(Az + nb + IC * p.px); pf_l2(Ax + nb + IC * p.px); }
    if (edge) { pf_l2(Ay + nb + (GROUP == 0 ? 4 : -1)); pf_l2(Az + nb + (GROUP == 0 ? 4 : -1)); }
    pf_l2(p.F[0] + nb); pf_l2(p.F[1] + nb); pf_l2(p.F[2] + nb);
    const long long nm = p.mplane * (long long)(zp - 1) + mo;
    if constexpr (MARR == 1) {
      if (!m_uniform) { pf_l2(p.m_arr[0] + nm); pf_l2(p.m_arr[1] + nm); pf_l2(p.m_arr[2] + nm); }
    }
    if constexpr (GENERAL) {
      if (hasx) {
        const long long xsl = (long long)p.cxp * p.n[1] * (long long)(zp - 1) + xs_off;
        pf_l2(p.W[0] + xsl); pf_l2(p.U[2] + xsl);
      }
      if (hasy) {
        const long long ysl = (long long)p.mpx * p.cy * (long long)(zp - 1) + ys_off;
        pf_l2(p.W[1] + ysl); pf_l2(p.U[0] + ysl);
      }
      if (PZA && p.sg[2][zp - 1] != T(0)) {
        const long long zsn = p.mplane * (long long)p.slab[2].idx(zp) + mo;
        pf_l2(p.W[2] + zsn); pf_l2(p.U[1] + zsn);
      }
    }
    if constexpr (EXTRAS) {
      // the pole / conductivity arrays are read through sigma-dependent chains: make them L2 hits
      if constexpr (GROUP == 1) {
        for (int q = 0; q < p.npole; ++q) {
          const PoleDesc<T>& pl = p.pole_ext ? p.pole_ext[q] : p.pole[q];
          pf_l2(pl.sigma + nm);
#pragma unroll
          for (int d = 0; d < 3; ++d) { pf_l2(pl.Pc[d] + nm); pf_l2(pl.Pp[d] + nm); }
        }
        if (p.chi3 != nullptr) pf_l2(p.chi3 + nm);
        if ((p.npole > 0 || p.chi3 != nullptr) && p.Dst[0] != nullptr) { pf_l2(p.Dst[0] + nm); pf_l2(p.Dst[1] + nm); pf_l2(p.Dst[2] + nm); }
      }
      if (p.sigD[0] != nullptr) {
        pf_l2(p.sigD[0] + nm); pf_l2(p.sigD[1] + nm); pf_l2(p.sigD[2] + nm);
        if (p.C[0] != nullptr) { pf_l2(p.C[0] + nm); pf_l2(p.C[1] + nm); pf_l2(p.C[2] + nm); }
      }
    }
  };

  // ---- chain mode: wait until the other field group has finished the z chunks this tile reads
  // (E reads H[k-1..k]: chunks c-1, c; H reads E[k..k+1]: chunks c, c+1).  The same wait also
  // covers the write-after-read hazard on the own field (its readers are exactly those items).
  if (p.dep_on) {
    // ask L2 for the first planes before blocking on the counters (stale lines cannot result:
    // L2 is the coherence point, the producers' stores update the prefetched lines)
    if (act) {
      const long long cb = p.plane * (long long)(GROUP == 0 ? it.z0 : it.z0 - 1) + fo;
      pf_l2(Ax + cb); pf_l2(Ay + cb);
      prefetch_plane(it.z0);
    }
    if (threadIdx.x < 2) {
      // the two chunk counters are polled by two lanes at once (one L2 round trip instead of two)
      const int c = it.chunk + (GROUP == 0 ? 0 : -1) + (int)threadIdx.x;
      if (c >= 0 && c < p.nchunk) {
        const unsigned long long target = p.epoch_other * p.cnt_other[c];
        unsigned spins = 0;
        while (ld_acquire_u64(p.done_other + c) < target) {
          __nanosleep(40);
          // never hang the GPU: give up after ~1 s, or at once when another CTA already gave up
          if (++spins > (1u << 20) || ((spins & 1023u) == 0 && *(volatile int*)p.err_flag != 0)) { *p.err_flag = 1; break; }
        }
      }
    }
    __syncthreads();
  }

  // ---- z-neighbour carry ----
  V4<T> ax_c = zero4<T>(), ay_c = zero4<T>();  // GROUP 0: current plane; GROUP 1: plane below
  {
    const long long b0 = p.plane * (long long)(GROUP == 0 ? it.z0 : it.z0 - 1) + fo;
    if (act) { ax_c = ld4(Ax + b0); ay_c = ld4(Ay + b0); }
  }

  for (int iz = it.z0; iz < z_end; ++iz) {
    const long long base = p.plane * (long long)iz + fo;
    const long long mbase = p.mplane * (long long)(iz - 1) + mo;
    const bool more = iz + 1 < z_end;
    if constexpr (NU) idz_ = p.idv[2][iz - 1];
#if KHR_PREFETCH
    if (act) {
      if (iz == it.z0) {
        for (int zp = iz + 1; zp < iz + KHR_PF_DIST && zp < z_end; ++zp) prefetch_plane(zp);
      }
      if (iz + KHR_PF_DIST < z_end) prefetch_plane(iz + KHR_PF_DIST);
    }
#endif
    V4<T> ax0, ay0, az0, ax_z, ay_z, az_y, ax_y;
    V4<T> fx, fy, fz, m0, m1, m2;
    T ay_x = T(0), az_x = T(0);
    T* __restrict__ Fx = p.F[0] + base;
    T* __restrict__ Fy = p.F[1] + base;
    T* __restrict__ Fz = p.F[2] + base;
    // ---- general path: plane coefficients, aux addresses, aux loads (all issued up front) ----
    Co<T> czc;
    bool hasz = false;
    V4<T> ux, uy, uz, wx, wy, wz;
    T *Uxp = nullptr, *Uyp = nullptr, *Uzp = nullptr, *Wxp = nullptr, *Wyp = nullptr, *Wzp = nullptr;
    if constexpr (GENERAL) {
      czc = czn;
      if constexpr (PZA) {
        if (more) { czn.s = p.sg[2][iz]; czn.om = p.om[2][iz]; czn.ip = p.ip[2][iz]; }
        hasz = act && (czc.s != T(0));
      }
      ux = uy = uz = wx = wy = wz = zero4<T>();
      if (hasx) {
        const long long xst = (long long)p.cxp * p.n[1];
        const long long xsl = xst * (long long)(iz - 1) + xs_off;
        Wxp = p.W[0] + xsl; Uzp = p.U[2] + xsl;
        wx = ld4(Wxp); uz = ld4(Uzp);
      }
      if (hasy) {
        const long long yst = (long long)p.mpx * p.cy;
        const long long ysl = yst * (long long)(iz - 1) + ys_off;
        Wyp = p.W[1] + ysl; Uxp = p.U[0] + ysl;
        wy = ld4(Wyp); ux = ld4(Uxp);
      }
      if (hasz) {
        const long long zsl = p.mplane * (long long)p.slab[2].idx(iz) + mo;
        Wzp = p.W[2] + zsl; Uyp = p.U[1] + zsl;
        wz = ld4(Wzp); uy = ld4(Uyp);
      }
    }
    if (act) {
      if constexpr (GROUP == 0) {
        ax0 = ax_c; ay0 = ay_c;
        ax_z = ld4(Ax + base + p.plane);
        ay_z = ld4(Ay + base + p.plane);
      } else {
        ax_z = ax_c; ay_z = ay_c;
        ax0 = ld4(Ax + base);
        ay0 = ld4(Ay + base);
      }
      az0 = ld4(Az + base);
      az_y = ld4(Az + base + IC * p.px);
      ax_y = ld4(Ax + base + IC * p.px);
      if (edge) {
        ay_x = Ay[base + (GROUP == 0 ? 4 : -1)];
        az_x = Az[base + (GROUP == 0 ? 4 : -1)];
      }
      fx = ld4(Fx); fy = ld4(Fy); fz = ld4(Fz);
      if constexpr (MARR == 1) {
        if (m_uniform) {
#pragma unroll
          for (int e = 0; e < 4; ++e) { m0.v[e] = (T)it.mu[0]; m1.v[e] = (T)it.mu[1]; m2.v[e] = (T)it.mu[2]; }
        } else {
          m0 = ld4(p.m_arr[0] + mbase); m1 = ld4(p.m_arr[1] + mbase); m2 = ld4(p.m_arr[2] + mbase);
        }
      }
    } else {
      ax0 = ay0 = az0 = ax_z = ay_z = az_y = ax_y = zero4<T>();
    }
#define KHR_M0(e) (MARR == 1 ? m0.v[e] : (MARR == 2 ? mu0 : p.m_inv))
#define KHR_M1(e) (MARR == 1 ? m1.v[e] : (MARR == 2 ? mu1 : p.m_inv))
#define KHR_M2(e) (MARR == 1 ? m2.v[e] : (MARR == 2 ? mu2 : p.m_inv))
    // x neighbour by shuffle inside the LX-lane row
    {
      T sy, sz;
      if constexpr (GROUP == 0) {
        sy = __shfl_down_sync(0xffffffffu, ay0.v[0], 1, LX);
        sz = __shfl_down_sync(0xffffffffu, az0.v[0], 1, LX);
      } else {
        sy = __shfl_up_sync(0xffffffffu, ay0.v[3], 1, LX);
        sz = __shfl_up_sync(0xffffffffu, az0.v[3], 1, LX);
      }
      if (!edge) { ay_x = sy; az_x = sz; }
    }
    if (act) {
      T kx[4], ky[4], kz[4];
      T ku[3][4];  // unscaled K (only live in the full kernel)
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        T ayx, azx;
        if constexpr (GROUP == 0) {
          ayx = (e < 3) ? ay0.v[(e + 1) & 3] : ay_x;
          azx = (e < 3) ? az0.v[(e + 1) & 3] : az_x;
        } else {
          ayx = (e > 0) ? ay0.v[(e + 3) & 3] : ay_x;
          azx = (e > 0) ? az0.v[(e + 3) & 3] : az_x;
        }
        // K = dt * curl (Helpers.jl:286-298), same operation order as the reference
        const T k0 = dt * (idz_ * (ay_z.v[e] - ay0.v[e]) - idy_ * (az_y.v[e] - az0.v[e]));
        const T idxe = NU ? idxv[e] : idx_;
        const T k1 = dt * (idxe * (azx - az0.v[e]) - idz_ * (ax_z.v[e] - ax0.v[e]));
        const T k2 = dt * (idy_ * (ax_y.v[e] - ax0.v[e]) - idxe * (ayx - ay0.v[e]));
        kx[e] = KHR_M0(e) * k0;
        ky[e] = KHR_M1(e) * k1;
        kz[e] = KHR_M2(e) * k2;
        if constexpr (EXTRAS) { ku[0][e] = k0; ku[1][e] = k1; ku[2][e] = k2; }
      }

      if constexpr (!GENERAL) {
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          if (valid[e]) { fx.v[e] = fx.v[e] + kx[e]; fy.v[e] = fy.v[e] + ky[e]; fz.v[e] = fz.v[e] + kz[e]; }
        }
      } else {
        // ---- general path ----
        Co<T> cyv[4], czv[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) { cyv[e] = cyc; czv[e] = czc; }
        T so[3][4], sn[3][4], sd[3][4];
#pragma unroll
        for (int d = 0; d < 3; ++d)
#pragma unroll
          for (int e = 0; e < 4; ++e) { so[d][e] = T(0); sn[d][e] = T(0); sd[d][e] = T(0); }
        V4<T> c0 = zero4<T>(), c1 = zero4<T>(), c2 = zero4<T>();
        bool has_sd = false, use_c = false;
        unsigned polmask = 0;      // bit q: pole q is present in one of the thread's 4 cells
        bool pol_any = false;
        bool disp[4] = {false, false, false, false};  // cell carries a pole or a Kerr coefficient: D is kept
        V4<T> c3 = zero4<T>();                         // chi3 of the 4 cells
        bool chi_any = false;
        int sslot[3][4];                               // flux-accumulator slot of a source voxel, -1 if none
        T su[3][4], pu[3][4];                          // unscaled S(t) and sum of P^n
#pragma unroll
        for (int d = 0; d < 3; ++d)
#pragma unroll
          for (int e = 0; e < 4; ++e) { su[d][e] = T(0); pu[d][e] = T(0); sslot[d][e] = -1; }
        if constexpr (EXTRAS) {
          // sources (Sources.jl:355-356): S = real(a(t) * A[x])
          if (srcmask != 0) {
            for (int q = 0; q < p.nsrc; ++q) {
              if (!((srcmask >> (q < 31 ? q : 31)) & 1u)) continue;
              const SrcDesc<T>& s = p.src_ext ? p.src_ext[q] : p.src[q];
              const int ly = iy - s.s[1], lz = iz - s.s[2];
              if (ly < 0 || ly >= s.d[1] || lz < 0 || lz >= s.d[2]) continue;
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                const int lx = gx + e - s.s[0];
                if (lx >= 0 && lx < s.d[0]) {
                  const size_t ai = 2 * ((size_t)lx + (size_t)s.d[0] * ((size_t)ly + (size_t)s.d[1] * (size_t)lz));
                  const T are = s.amp[ai], aim = s.amp[ai + 1];
                  const int sl = s.slot[ai >> 1];
                  const T vn = s.an_re * are - s.an_im * aim;
                  const T vo = s.ao_re * are - s.ao_im * aim;
                  if (s.comp == 0) { sn[0][e] += KHR_M0(e) * vn; so[0][e] += KHR_M0(e) * vo; su[0][e] += vn; sslot[0][e] = sl; }
                  else if (s.comp == 1) { sn[1][e] += KHR_M1(e) * vn; so[1][e] += KHR_M1(e) * vo; su[1][e] += vn; sslot[1][e] = sl; }
                  else { sn[2][e] += KHR_M2(e) * vn; so[2][e] += KHR_M2(e) * vo; su[2][e] += vn; sslot[2][e] = sl; }
                }
              }
            }
          }
          // polarisation: the stored E carries -eps^-1 P^{n-1}; this step puts -eps^-1 P^n
          if constexpr (GROUP == 1) {
            if (p.chi3 != nullptr) {
              c3 = ld4(p.chi3 + mbase);
#pragma unroll
              for (int e = 0; e < 4; ++e) { disp[e] |= (c3.v[e] != T(0)); chi_any |= (c3.v[e] != T(0)); }
            }
            for (int q = 0; q < p.npole; ++q) {
              const PoleDesc<T>& pl = p.pole_ext ? p.pole_ext[q] : p.pole[q];
              V4<T> sg = ld4(pl.sigma + mbase);
              if ((sg.v[0] != T(0)) || (sg.v[1] != T(0)) || (sg.v[2] != T(0)) || (sg.v[3] != T(0))) {
                polmask |= 1u << q;
                pol_any = true;
#pragma unroll
                for (int e = 0; e < 4; ++e) disp[e] |= (sg.v[e] != T(0));
#pragma unroll
                for (int d = 0; d < 3; ++d) {
                  V4<T> pc = ld4(pl.Pc[d] + mbase), pp = ld4(pl.Pp[d] + mbase);
#pragma unroll
                  for (int e = 0; e < 4; ++e) {
                    const T mm = (d == 0) ? KHR_M0(e) : (d == 1) ? KHR_M1(e) : KHR_M2(e);
                    sn[d][e] -= mm * pc.v[e];
                    so[d][e] -= mm * pp.v[e];
                    pu[d][e] += pc.v[e];
                  }
                }
              }
            }
          }
          // material conductivity
          has_sd = p.sigD[0] != nullptr;
          if (has_sd) {
            bool anysd = false;
#pragma unroll
            for (int d = 0; d < 3; ++d) {
              V4<T> s4 = ld4(p.sigD[d] + mbase);
#pragma unroll
              for (int e = 0; e < 4; ++e) { sd[d][e] = T(0.5) * (dt * s4.v[e]); anysd |= (s4.v[e] != T(0)); }
            }
            use_c = anysd && (p.C[0] != nullptr) && (hasx || hasy || hasz);
            if (use_c) { c0 = ld4(p.C[0] + mbase); c1 = ld4(p.C[1] + mbase); c2 = ld4(p.C[2] + mbase); }
          }
        }
        // x: next = y, prev = z, own = x
        cascade<T, EXTRAS>(kx, cyv, czv, cx, hasy, hasx, ux, wx, fx, so[0], sn[0], has_sd, sd[0], c0, valid);
        // y: next = z, prev = x, own = y
        cascade<T, EXTRAS>(ky, czv, cx, cyv, hasz, hasy, uy, wy, fy, so[1], sn[1], has_sd, sd[1], c1, valid);
        // z: next = x, prev = y, own = z
        cascade<T, EXTRAS>(kz, cx, cyv, czv, hasx, hasz, uz, wz, fz, so[2], sn[2], has_sd, sd[2], c2, valid);
        if (hasx) { st4(Wxp, wx); st4(Uzp, uz); }
        if (hasy) { st4(Wyp, wy); st4(Uxp, ux); }
        if (hasz) { st4(Wzp, wz); st4(Uyp, uy); }
        if constexpr (EXTRAS) {
          // Source voxels outside the PML keep the flux field (B or D) like the reference and
          // rebuild A = m^-1 (T + S) from it every step (Helpers.jl:332-338).  With |S| >> |T| the
          // eliminated form would let the S round-off random-walk inside the stored field.
          if (srcmask != 0 && p.Tsrc != nullptr) {
#pragma unroll
            for (int d = 0; d < 3; ++d) {
              V4<T>& f = (d == 0) ? fx : (d == 1) ? fy : fz;
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                const bool in_pml = (cx[e].s != T(0)) || (cyc.s != T(0)) || (czc.s != T(0));
                if (sslot[d][e] >= 0 && !in_pml && !disp[e] && valid[e]) {
                  const T s_ = sd[d][e];
                  const T t_old = p.Tsrc[sslot[d][e]];
                  const T t_new = (s_ != T(0)) ? (((T(1) - s_) * t_old + ku[d][e]) / (T(1) + s_)) : (t_old + ku[d][e]);
                  T net = t_new;
                  net += su[d][e];
                  const T mm = (d == 0) ? KHR_M0(e) : (d == 1) ? KHR_M1(e) : KHR_M2(e);
                  f.v[e] = mm * net;
                  p.Tsrc[sslot[d][e]] = t_new;
                }
              }
            }
          }
        }
        if constexpr (EXTRAS && GROUP == 1) {
          // Dispersive voxels outside the PML keep D exactly like the reference
          // (Kernels.jl:315,371: fPD present -> no D elimination): D += K (or the sigma_D stage),
          // E = eps^-1 * ((D + S) - P) (Helpers.jl:332-338).  |P| >> |eps E| in a metal, so
          // recomputing E from D each step avoids a random walk of the P round-off in E.
          if ((pol_any || chi_any) && p.Dst[0] != nullptr) {
#pragma unroll
            for (int d = 0; d < 3; ++d) {
              V4<T> dv = ld4(p.Dst[d] + mbase);
              V4<T>& f = (d == 0) ? fx : (d == 1) ? fy : fz;
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                const bool in_pml = (cx[e].s != T(0)) || (cyc.s != T(0)) || (czc.s != T(0));
                if (disp[e] && !in_pml && valid[e]) {
                  const T s_ = sd[d][e];
                  const T d_new = (s_ != T(0)) ? (((T(1) - s_) * dv.v[e] + ku[d][e]) / (T(1) + s_)) : (dv.v[e] + ku[d][e]);
                  T net = d_new;
                  net += su[d][e];
                  net -= pu[d][e];
                  const T mm = (d == 0) ? KHR_M0(e) : (d == 1) ? KHR_M1(e) : KHR_M2(e);
                  f.v[e] = mm * net;
                  dv.v[e] = d_new;
                }
              }
              st4(p.Dst[d] + mbase, dv);
            }
          }
        }
        if constexpr (EXTRAS && GROUP == 1) {
          // Kerr correction (Dispersive.jl:127-148): E <- E / (1 + chi3 |E|^2), the three components
          // taken at the same array index, applied to the freshly rebuilt E = eps^-1 (D + S - P) and
          // before the ADE update, which therefore sees the corrected E (Kernels.jl:76-83)
          if (chi_any) {
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              if (c3.v[e] != T(0) && valid[e]) {
                const T ex = fx.v[e], ey = fy.v[e], ez = fz.v[e];
                const T e_sq = (ex * ex + ey * ey) + ez * ez;
                const T corr = T(1) / (T(1) + c3.v[e] * e_sq);
                fx.v[e] = ex * corr; fy.v[e] = ey * corr; fz.v[e] = ez * corr;
              }
            }
          }
        }
        if constexpr (EXTRAS) {
          if (use_c) { st4(p.C[0] + mbase, c0); st4(p.C[1] + mbase, c1); st4(p.C[2] + mbase, c2); }
          // ADE (Dispersive.jl:25-88): P^{n+1} from P^n, P^{n-1} and the new E; written over P^{n-1}
          if constexpr (GROUP == 1) {
            for (int q = 0; q < p.npole; ++q) {
              if ((polmask >> q) & 1u) {
                const PoleDesc<T>& pl = p.pole_ext ? p.pole_ext[q] : p.pole[q];
                V4<T> sg = ld4(pl.sigma + mbase);
#pragma unroll
                for (int d = 0; d < 3; ++d) {
                  V4<T> pc = ld4(pl.Pc[d] + mbase), pp = ld4(pl.Pp[d] + mbase);
                  const V4<T>& en = (d == 0) ? fx : (d == 1) ? fy : fz;
                  V4<T> pn;
#pragma unroll
                  for (int e = 0; e < 4; ++e) {
                    T v = pl.g1i * ((pl.cp * pc.v[e] - pl.g1 * pp.v[e]) + pl.cd * sg.v[e] * en.v[e]);
                    pn.v[e] = (sg.v[e] != T(0) && valid[e]) ? v : pc.v[e];
                  }
                  st4(pl.Pp[d] + mbase, pn);
                }
              }
            }
          }
        }
      }
      st4(Fx, fx);
      st4(Fy, fy);
      st4(Fz, fz);
    }
    // carry
    if constexpr (GROUP == 0) { ax_c = ax_z; ay_c = ay_z; }
    else { ax_c = ax0; ay_c = ay0; }
  }
  if (p.dep_on) {
    // release: every store of the CTA is ordered before the counter increment
    __syncthreads();
    if (threadIdx.x == 0) {
      __threadfence();
      atomicAdd(p.done_mine + it.chunk, 1ull);
    }
  }
}

template <class T, int GROUP, int MODE, int MARR, int AXM = 7, bool NU = false>
__global__ void __launch_bounds__(CTA, (min_ctas<T, MODE, AXM>())) step_kernel(const __grid_constant__ StepParams<T> p) {
  step_body<T, GROUP, MODE, MARR, AXM, NU>(p, p.items[blockIdx.x]);
}

// ----------------------------------------------------------------------------
// Sweep mode: ONE grid per time step holds the interior and PML tiles of BOTH half-steps in
// z-chunk-major order — H(chunk 0), E(chunk 0), H(chunk 1), E(chunk 1), ... — so that the E tiles
// of a chunk run right after its H tiles, while the H planes just written and the E planes just
// read are still in the 126 MB L2: the E half-step then takes its curl operand and its own field
// from L2 instead of HBM, and only one grid-wide drain per step remains.  Ordering is the chain
// mode's: an E tile waits (device counters) for the H tiles of chunks c-1, c of this step, an H
// tile for the E tiles of chunks c, c+1 of the previous step; every dependency points to a tile
// with a lower block index or an earlier launch, so in-order block dispatch guarantees progress.
// WorkItem::flags bit 8 = field group, bit 9 = PML tile.
// ----------------------------------------------------------------------------
template <class T, int MH, int ME>
__global__ void __launch_bounds__(CTA, (min_ctas<T, 1, 7>())) sweep_kernel(const __grid_constant__ StepParams<T> ph,
                                                                         const __grid_constant__ StepParams<T> pe,
                                                                         const WorkItem* __restrict__ items) {
  const WorkItem it = items[blockIdx.x];
  const int sel = (it.flags >> 8) & 3;   // block-uniform
  if (sel == 0) step_body<T, 0, 0, MH>(ph, it);
  else if (sel == 2) step_body<T, 0, 1, MH>(ph, it);
  else if (sel == 1) step_body<T, 1, 0, ME>(pe, it);
  else step_body<T, 1, 1, ME>(pe, it);
}

// ----------------------------------------------------------------------------
// Fix-up of the source / dispersive / Kerr voxels OUTSIDE the PML (round 2).  The marching kernels run such tiles as
// plain interior tiles (MODE 0, or the TMA kernel) and this kernel then rewrites the affected voxels with exactly the
// arithmetic of the MODE 2 path for them, which never used the marched value there anyway: the flux field is kept on
// those voxels like the reference keeps D / B (Kernels.jl:315,371; Helpers.jl:332-338), so
//     D <- D + K,   E = eps^-1 ((D + S) - P),   Kerr correction (Dispersive.jl:127-148),   ADE update (Dispersive.jl:25-88)
// need only the curl operand (untouched during the half-step), D / the source flux slot, S and P — not the marched E.
// One thread per voxel, high occupancy; the latency-bound 250-register MODE 2 launch then only
// remains for tiles inside the PML and for conductive media.  Voxels without a source or pole are left alone.
// ----------------------------------------------------------------------------
template <class T, int GROUP, int MARR, bool NU>
__global__ void __launch_bounds__(256) fixup_kernel(const __grid_constant__ StepParams<T> p) {
  constexpr int IC = (GROUP == 0) ? 1 : -1;
  const WorkItem it = p.items[blockIdx.x];
  const int ncell = it.xw * it.yh * it.zn;
  const T* __restrict__ A0 = p.A[0];
  const T* __restrict__ A1 = p.A[1];
  const T* __restrict__ A2 = p.A[2];
  // one voxel per thread, blockIdx.y walks the tile in slices of 256 voxels: enough independent threads in flight to
  // hide the dependent loads (a per-thread loop over the tile measured 5x slower, profiles/r02_mode2_ab.txt)
  for (int q = blockIdx.y * blockDim.x + threadIdx.x; q < ncell; q += gridDim.y * blockDim.x) {
    const int ix = it.x0 + q % it.xw, iy = it.y0 + (q / it.xw) % it.yh, iz = it.z0 + q / (it.xw * it.yh);
    const long long fo = p.plane * (long long)iz + (long long)p.px * iy + (ix + XO);
    const long long mo = p.mplane * (long long)(iz - 1) + (long long)p.mpx * (iy - 1) + (ix - 1);
    // ---- what lives on this voxel ----
    T su[3] = {T(0), T(0), T(0)}, pu[3] = {T(0), T(0), T(0)};
    int sslot[3] = {-1, -1, -1};
    bool disp = false;
    if (it.flags & 1) {
      for (int s_ = 0; s_ < p.nsrc; ++s_) {
        const SrcDesc<T>& s = p.src_ext ? p.src_ext[s_] : p.src[s_];
        const int lx = ix - s.s[0], ly = iy - s.s[1], lz = iz - s.s[2];
        if (lx < 0 || lx >= s.d[0] || ly < 0 || ly >= s.d[1] || lz < 0 || lz >= s.d[2]) continue;
        const size_t ai = 2 * ((size_t)lx + (size_t)s.d[0] * ((size_t)ly + (size_t)s.d[1] * (size_t)lz));
        const T are = s.amp[ai], aim = s.amp[ai + 1];
        su[s.comp] += s.an_re * are - s.an_im * aim;      // S = real(a(t) A[x]) (Sources.jl:355-356)
        sslot[s.comp] = s.slot[ai >> 1];
      }
    }
    T c3 = T(0);
    unsigned polmask = 0;
    if constexpr (GROUP == 1) {
      if (p.chi3 != nullptr) { c3 = p.chi3[mo]; disp |= (c3 != T(0)); }
      for (int k = 0; k < p.npole; ++k) {
        const PoleDesc<T>& pl = p.pole_ext ? p.pole_ext[k] : p.pole[k];
        if (pl.sigma[mo] != T(0)) {
          polmask |= 1u << k;
          disp = true;
#pragma unroll
          for (int d = 0; d < 3; ++d) pu[d] += pl.Pc[d][mo];
        }
      }
    }
    if (!disp && sslot[0] < 0 && sslot[1] < 0 && sslot[2] < 0) continue;
    // ---- K = dt * curl, the expression and operation order of step_body ----
    const T dt = p.dt;
    T idx_ = p.idl[0], idy_ = p.idl[1], idz_ = p.idl[2];
    if constexpr (NU) { idx_ = p.idv[0][ix - 1]; idy_ = p.idv[1][iy - 1]; idz_ = p.idv[2][iz - 1]; }
    const T ax0 = A0[fo], ay0 = A1[fo], az0 = A2[fo];
    const T ay_z = A1[fo + IC * p.plane], ax_z = A0[fo + IC * p.plane];
    const T az_y = A2[fo + IC * p.px], ax_y = A0[fo + IC * p.px];
    const T azx = A2[fo + IC], ayx = A1[fo + IC];
    T ku[3];
    ku[0] = dt * (idz_ * (ay_z - ay0) - idy_ * (az_y - az0));
    ku[1] = dt * (idx_ * (azx - az0) - idz_ * (ax_z - ax0));
    ku[2] = dt * (idy_ * (ax_y - ax0) - idx_ * (ayx - ay0));
    T f[3];
    bool touched[3] = {false, false, false};
#pragma unroll
    for (int d = 0; d < 3; ++d) {
      T mm;
      if constexpr (MARR == 1) mm = (it.flags & 2) ? (T)it.mu[d] : p.m_arr[d][mo];
      else mm = p.m_inv;
      if constexpr (GROUP == 1) {
        if (disp) {
          const T d_new = p.Dst[d][mo] + ku[d];
          T net = d_new;
          net += su[d];
          net -= pu[d];
          f[d] = mm * net;
          p.Dst[d][mo] = d_new;
          touched[d] = true;
          continue;
        }
      }
      if (sslot[d] >= 0) {
        const T t_new = p.Tsrc[sslot[d]] + ku[d];
        T net = t_new;
        net += su[d];
        f[d] = mm * net;
        p.Tsrc[sslot[d]] = t_new;
        touched[d] = true;
      }
    }
    if constexpr (GROUP == 1) {
      if (c3 != T(0)) {   // Kerr: the three components at the same array index (Dispersive.jl:127-148); disp => all touched
        const T e_sq = (f[0] * f[0] + f[1] * f[1]) + f[2] * f[2];
        const T corr = T(1) / (T(1) + c3 * e_sq);
        f[0] = f[0] * corr; f[1] = f[1] * corr; f[2] = f[2] * corr;
      }
    }
#pragma unroll
    for (int d = 0; d < 3; ++d)
      if (touched[d]) p.F[d][fo] = f[d];
    if constexpr (GROUP == 1) {
      for (int k = 0; k < p.npole; ++k) {
        if (!((polmask >> k) & 1u)) continue;
        const PoleDesc<T>& pl = p.pole_ext ? p.pole_ext[k] : p.pole[k];
        const T sg = pl.sigma[mo];
#pragma unroll
        for (int d = 0; d < 3; ++d) {
          const T pc = pl.Pc[d][mo], pp = pl.Pp[d][mo];
          pl.Pp[d][mo] = pl.g1i * ((pl.cp * pc - pl.g1 * pp) + pl.cd * sg * f[d]);
        }
      }
    }
  }
}

// ----------------------------------------------------------------------------
// Planner helper (runs once in khr_finalize_plan): marks the work items over which all three
// per-voxel constitutive arrays are bit-wise constant, and records the values.  Such tiles
// (most of a piecewise-homogeneous scene) skip the three material loads per plane; the
// arithmetic is unchanged, so results are bit-identical to the per-voxel path.
// ----------------------------------------------------------------------------
template <class T>
__global__ void __launch_bounds__(256) classify_items_kernel(WorkItem* items, const T* m0, const T* m1, const T* m2,
                                                            int mpx, long long mplane) {
  WorkItem& it = items[blockIdx.x];
  const T* arr[3] = {m0, m1, m2};
  const long long first = mplane * (long long)(it.z0 - 1) + (long long)mpx * (it.y0 - 1) + (it.x0 - 1);
  bool same = true;
  const int ncell = it.xw * it.yh * it.zn;
  for (int c = 0; c < 3; ++c) {
    const T v0 = arr[c][first];
    for (int q = threadIdx.x; q < ncell; q += blockDim.x) {
      const int x = q % it.xw, y = (q / it.xw) % it.yh, z = q / (it.xw * it.yh);
      const T v = arr[c][first + x + (long long)mpx * y + mplane * z];
      if (sizeof(T) == 4) same &= (__float_as_uint((float)v) == __float_as_uint((float)v0));
      else same &= (__double_as_longlong((double)v) == __double_as_longlong((double)v0));
    }
  }
  const int all = __syncthreads_and(same ? 1 : 0);
  if (threadIdx.x == 0 && all) {
    it.flags |= 2;
    for (int c = 0; c < 3; ++c) it.mu[c] = (double)arr[c][first];
  }
}

// ----------------------------------------------------------------------------
// Running DFT (Monitors.jl:331-379): M[x,y,z,k] += (dt * exp(i f_k 2 pi t)) * F[x+off]
// One launch covers every monitor of a field group that is due this step.
// ----------------------------------------------------------------------------
template <class T>
struct MonDesc {
  T* M;               // complex interleaved (nx,ny,nz,nf)
  const T* F;         // field array (ghosted layout)
  const T* Fi;        // imaginary part of the field (complex fields, Bloch boundaries) or null
  const T* freqs;     // nf values
  int s[3];           // local start cell (may start below 1 in z when clipped by host)
  int n[3];           // extent of the part this rank accumulates
  int moff[3];        // offset of that part inside the monitor box
  int mn[3];          // full monitor box extent (for M strides)
  int nf;
  int decimation;
  int group;
};

constexpr int DFT_SLOTS = 32;    // (monitor, frequency-range) slots per launch
constexpr int DFT_PHASORS = 384; // phasors per launch, passed by value with the launch

// The phasors dt*exp(i*T(f_k)*T(2 pi t)) are evaluated on the host exactly as the reference
// does (phase formed in T, Monitors.jl:323,355; sincos to < 1 ulp like Julia's) and travel in
// the kernel-parameter buffer, so no device table has to be kept alive or synchronised.
template <class T>
struct DftBatch {
  int n;
  int mon[DFT_SLOTS];     // monitor index
  int k0[DFT_SLOTS];      // first frequency of the slot
  int kc[DFT_SLOTS];      // number of frequencies
  int off[DFT_SLOTS];     // offset into ph_re / ph_im
  T ph_re[DFT_PHASORS];
  T ph_im[DFT_PHASORS];
};

template <class T>
__global__ void __launch_bounds__(256) dft_kernel(const MonDesc<T>* __restrict__ mons,
                                                  const __grid_constant__ DftBatch<T> bt, long long plane, int px) {
  const int slot = blockIdx.y;
  const MonDesc<T> m = mons[bt.mon[slot]];
  const long long ncell = (long long)m.n[0] * m.n[1] * m.n[2];
  const long long cell = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (cell >= ncell) return;
  const int x = (int)(cell % m.n[0]);
  const int y = (int)((cell / m.n[0]) % m.n[1]);
  const int z = (int)(cell / ((long long)m.n[0] * m.n[1]));
  const long long fo = plane * (long long)(m.s[2] + z) + (long long)px * (m.s[1] + y) + (m.s[0] + x + XO);
  const T f = m.F[fo];
  const long long mcell = (long long)(x + m.moff[0]) +
                          (long long)m.mn[0] * ((long long)(y + m.moff[1]) + (long long)m.mn[1] * (z + m.moff[2]));
  const long long mstride = (long long)m.mn[0] * m.mn[1] * m.mn[2];
  const int k0 = bt.k0[slot], kc = bt.kc[slot], off = bt.off[slot];
  for (int k = 0; k < kc; ++k) {
    T* mp = m.M + 2 * (mcell + mstride * (k0 + k));
    // M += (dt * e) * F  (complex accumulate, Monitors.jl:353-357)
    T re = mp[0], im = mp[1];
    if (m.Fi == nullptr) {
      mp[0] = re + bt.ph_re[off + k] * f;
      mp[1] = im + bt.ph_im[off + k] * f;
    } else {
      // complex field: (dt e) * (f + i fi) = (pr f - pi fi) + i (pr fi + pi f)
      const T fi = m.Fi[fo];
      const T pr = bt.ph_re[off + k], pi = bt.ph_im[off + k];
      mp[0] = re + (pr * f - pi * fi);
      mp[1] = im + (pr * fi + pi * f);
    }
  }
}

// ----------------------------------------------------------------------------
// Periodic wrap-around of one axis (Chunking.jl:1725-1770, single-chunk form): for the three
// components of the field group just updated, last interior layer N -> ghost 0 and first
// interior layer 1 -> ghost N+1, over the transverse cells 1..N (send / recv ranges of
// Chunking.jl:1825-1852).  `sa` is the element stride of the wrapped axis, `s1`, `s2` those
// of the transverse axes; `base` the offset of cell (0,0,0).
// ----------------------------------------------------------------------------
template <class T>
__global__ void __launch_bounds__(256) wrap_kernel(T* f0, T* f1, T* f2, long long base, long long sa, long long s1,
                                                   long long s2, int n_axis, int n1, int n2) {
  const long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= (long long)n1 * n2) return;
  const int a = (int)(q % n1) + 1, b = (int)(q / n1) + 1;
  const long long o = base + s1 * a + s2 * b;
  T* f = blockIdx.y == 0 ? f0 : (blockIdx.y == 1 ? f1 : f2);
  f[o] = f[o + sa * n_axis];              // ghost 0    <- cell N
  f[o + sa * (n_axis + 1)] = f[o + sa];   // ghost N+1  <- cell 1
}

// Bloch / periodic wrap-around of complex fields kept as two real arrays (real and imaginary
// parts, identical layout): copy, then multiply by the phase in ComplexF64 and store back as
// Complex{T} (Chunking.jl:1735-1764, 2163-2167).  rev = exp(-i k L) goes with the lower ghost,
// fwd = exp(+i k L) with the upper one; a factor of exactly 1 is skipped like the reference does.
template <class T>
struct BlochWrapArgs {
  T* fr[3];
  T* fi[3];
  long long base, sa, s1, s2;
  int n_axis, n1, n2;
  double rev_re, rev_im, fwd_re, fwd_im;
  int apply_rev, apply_fwd;
};
template <class T>
__global__ void __launch_bounds__(256) bloch_wrap_kernel(const __grid_constant__ BlochWrapArgs<T> a) {
  const long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= (long long)a.n1 * a.n2) return;
  const int i1 = (int)(q % a.n1) + 1, i2 = (int)(q / a.n1) + 1;
  const long long o = a.base + a.s1 * i1 + a.s2 * i2;
  T* fr = a.fr[blockIdx.y];
  T* fi = a.fi[blockIdx.y];
  {  // ghost 0 <- cell N
    double vr = (double)fr[o + a.sa * a.n_axis], vi = (double)fi[o + a.sa * a.n_axis];
    if (a.apply_rev) { const double r = vr * a.rev_re - vi * a.rev_im, i = vr * a.rev_im + vi * a.rev_re; vr = r; vi = i; }
    fr[o] = (T)vr; fi[o] = (T)vi;
  }
  {  // ghost N+1 <- cell 1
    double vr = (double)fr[o + a.sa], vi = (double)fi[o + a.sa];
    if (a.apply_fwd) { const double r = vr * a.fwd_re - vi * a.fwd_im, i = vr * a.fwd_im + vi * a.fwd_re; vr = r; vi = i; }
    fr[o + a.sa * (a.n_axis + 1)] = (T)vr; fi[o + a.sa * (a.n_axis + 1)] = (T)vi;
  }
}

// phase factor on one received ghost plane of a complex field (z ring over several ranks): (fr + i fi) *= (pr + i pi),
// the product in ComplexF64 and the result stored as Complex{T} like bloch_wrap_kernel
template <class T>
__global__ void __launch_bounds__(256) bloch_phase_kernel(T* __restrict__ fr, T* __restrict__ fi, long long n, double pr, double pi) {
  const long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= n) return;
  const double vr = (double)fr[q], vi = (double)fi[q];
  fr[q] = (T)(vr * pr - vi * pi);
  fi[q] = (T)(vr * pi + vi * pr);
}

// ----------------------------------------------------------------------------
// Poynting flux through a monitor plane (FluxMonitor.jl:92-156 get_flux), on the device:
//   S(f) = sum_cells real(E1 conj(H2) - E2 conj(H1)) * dA
// with the reference's arithmetic per cell (two-plane average and the complex products in
// Complex{T}, the area factor and the sum in Float64).  Deterministic: per-block partial
// sums, then one thread per frequency adds the partials in order.
// ----------------------------------------------------------------------------
template <class T>
struct FluxArgs {
  const T* M[4];     // e1, e2, h1, h2 accumulators, complex interleaved (nx,ny,nz,nf)
  int n[4][3];       // their box extents
  int normal, t1, t2, n1, n2, nf;
  double dA;
};

template <class T>
__device__ __forceinline__ void flux_val(const FluxArgs<T>& a, int m, int i1, int i2, int kf, T& re, T& im) {
  const int* n = a.n[m];
  const size_t ncell = (size_t)n[0] * n[1] * n[2];
  int idx[3];
  idx[a.t1] = i1; idx[a.t2] = i2; idx[a.normal] = 0;
  const size_t c0 = (size_t)kf * ncell + (size_t)idx[0] + (size_t)n[0] * ((size_t)idx[1] + (size_t)n[1] * idx[2]);
  re = a.M[m][2 * c0]; im = a.M[m][2 * c0 + 1];
  if (n[a.normal] >= 2) {  // _avg_dim: (f[1] + f[2]) / 2 in Complex{T}
    idx[a.normal] = 1;
    const size_t c1 = (size_t)kf * ncell + (size_t)idx[0] + (size_t)n[0] * ((size_t)idx[1] + (size_t)n[1] * idx[2]);
    re = (re + a.M[m][2 * c1]) / T(2);
    im = (im + a.M[m][2 * c1 + 1]) / T(2);
  }
}

template <class T>
__global__ void __launch_bounds__(256) flux_kernel(const __grid_constant__ FluxArgs<T> a, double* __restrict__ partial) {
  const int kf = blockIdx.y;
  const long long ncell = (long long)a.n1 * a.n2;
  double s = 0.0;
  for (long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x; q < ncell; q += (long long)gridDim.x * blockDim.x) {
    const int i1 = (int)(q % a.n1), i2 = (int)(q / a.n1);
    T e1r, e1i, e2r, e2i, h1r, h1i, h2r, h2i;
    flux_val(a, 0, i1, i2, kf, e1r, e1i); flux_val(a, 1, i1, i2, kf, e2r, e2i);
    flux_val(a, 2, i1, i2, kf, h1r, h1i); flux_val(a, 3, i1, i2, kf, h2r, h2i);
    const T re1 = e1r * h2r + e1i * h2i;   // real(et1 * conj(ht2))
    const T re2 = e2r * h1r + e2i * h1i;   // real(et2 * conj(ht1))
    s += (double)(re1 - re2) * a.dA;
  }
  for (int o = 16; o > 0; o >>= 1) s += __shfl_down_sync(0xffffffffu, s, o);
  __shared__ double ws[8];
  if ((threadIdx.x & 31) == 0) ws[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0;
    for (int q = 0; q < 8; ++q) t += ws[q];
    partial[(size_t)kf * gridDim.x + blockIdx.x] = t;
  }
}

__global__ void flux_finish_kernel(const double* __restrict__ partial, int nblocks, int nf, double* __restrict__ out) {
  const int kf = blockIdx.x * blockDim.x + threadIdx.x;
  if (kf >= nf) return;
  double t = 0;
  for (int q = 0; q < nblocks; ++q) t += partial[(size_t)kf * nblocks + q];
  out[kf] = t;
}

// sum of squares (Simulation.jl:440-445 stop_when_dft_decayed reduces |M|^2 every step).
// Deterministic like the flux reduction: fixed grid-stride assignment, one partial per block,
// then one thread per monitor adds the partials in block order (no atomics).
template <class T>
__global__ void __launch_bounds__(256) sumsq_kernel(const T* __restrict__ a, long long n, double* __restrict__ partial) {
  double s = 0;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    double v = (double)a[i];
    s += v * v;
  }
  for (int o = 16; o > 0; o >>= 1) s += __shfl_down_sync(0xffffffffu, s, o);
  __shared__ double ws[8];
  if ((threadIdx.x & 31) == 0) ws[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0;
    for (int q = 0; q < 8; ++q) t += ws[q];
    partial[blockIdx.x] = t;
  }
}
// out[m] = sum of partial[m * stride .. + nblocks[m]) in order; nblocks[m] == 0 leaves out[m] untouched
__global__ void sumsq_finish_kernel(const double* __restrict__ partial, const int* __restrict__ nblocks, int stride, int nmon,
                                    double* __restrict__ out) {
  const int m = blockIdx.x * blockDim.x + threadIdx.x;
  if (m >= nmon || nblocks[m] == 0) return;
  double t = 0;
  for (int q = 0; q < nblocks[m]; ++q) t += partial[(size_t)m * stride + q];
  out[m] = t;
}

}  // namespace khr
