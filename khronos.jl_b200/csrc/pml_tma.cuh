// ============================================================================
// pml_tma.cuh — the PML half-step as a persistent, TMA-fed plane pipeline (sm_100a)
//
// Same arithmetic as step_kernel<T, GROUP, MODE = 1> (the C -> U -> T -> W cascade of
// Helpers.jl:30-33, 39-270, 323-366 on the B/D-eliminated form, `cascade<>` of
// step_kernels.cuh is reused verbatim, so the results are bit-identical), different data
// movement:
//
//   * one CTA per SM, persistent: 8 consumer warps + 1 producer warp; work items (x/y tile x
//     z segment) are claimed with an atomic counter, so there is no wave quantisation and no
//     per-tile ramp — the producer simply runs on into the next tile;
//   * the producer warp moves every array of a z plane of the tile — the three curl operands
//     with their +-1 halo row / column, the three updated fields, the U / W slabs of the PML
//     axes the tile lies in and the per-voxel material arrays — with `cp.async.bulk.tensor.3d`
//     (TMA, SASS UTMALDG) into a ring of shared-memory stages, completion on an mbarrier
//     (`complete_tx::bytes`); 4 stages of ~50 KB keep ~150-200 KB per SM in flight, which is
//     what hides the DRAM latency (the LDG version needed occupancy for that and had 16 warps);
//   * consumers read the stage with LDS.128 at compile-time offsets from one base register
//     (no 64-bit address arithmetic, no prefetch instructions, no shuffles: the x / y
//     neighbours are just other words of the staged box), run the cascade in registers and
//     write the results with STG.128;
//   * the z neighbour plane is carried in registers: the E half-step marches up (needs H at
//     k-1), the H half-step marches DOWN (needs E at k+1), so in both cases the plane loaded
//     for iteration i is the neighbour plane of iteration i+1; the first neighbour plane of a
//     tile and its per-column / per-row PML coefficients travel in a small per-tile ring entry.
//
// Uniform grids, Float32, MODE 1 tiles (no sources / conductivity / poles); everything else
// stays on step_kernel.
// ============================================================================
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "step_kernels.cuh"

namespace khr {

constexpr int TMA_CONSUMERS = 256;            // 8 consumer warps, thread = 4 x cells like step_kernel
constexpr int TMA_THREADS = TMA_CONSUMERS + 32;
constexpr int TMA_ABOX = 4864;                // bytes reserved for a halo box: max((32+4)*33, (64+4)*17, (128+4)*9) * 4 = 4752 -> 38 * 128
constexpr int TMA_OBOX = 4096;                // owner box: 1024 cells * 4
constexpr int TMA_HDR = 128;
constexpr int TMA_ENTRY = 2 * TMA_ABOX;        // per-tile ring entry: the neighbour planes of Ax, Ay the march starts from
constexpr int TMA_MAXSTAGES = 6;
// optional register cap of the persistent kernels (-DKHR_TMA_MAXREG=112: 9 warps x 112 registers would leave room on the SM
// for one 128-thread MODE 2 CTA).  Measured: no gain, 0.3-1 % slower (profiles/r02_tma_ab.txt r2c19) -> off
#ifndef KHR_TMA_MAXREG
#define KHR_TMA_MAXREG 0
#endif
#if KHR_TMA_MAXREG > 0
#define KHR_TMA_BOUNDS __maxnreg__(KHR_TMA_MAXREG)
#else
#define KHR_TMA_BOUNDS __launch_bounds__(TMA_THREADS, 1)
#endif

__host__ __device__ constexpr int tma_stage_bytes(int marr) { return TMA_HDR + 3 * TMA_ABOX + 3 * TMA_OBOX + 6 * TMA_OBOX + (marr == 1 ? 3 * TMA_OBOX : 0); }
__host__ __device__ constexpr int tma_smem_bytes(int marr, int stages) { return 1024 + 2 * TMA_ENTRY + stages * tma_stage_bytes(marr); }

struct TmaHdr {      // one per stage, written by the producer, 64 bytes
  int x0, y0, xw, yh;
  int iz, lx_log2, flags, entry;
  int q, z0, zn, cx0;          // plane number within the item, the item's z range, compact x index of x0
  float mu0, mu1, mu2;
  int chunk;                   // z chunk of the item (fused step: completion counters)
};
constexpr int TF_FIRST = 1, TF_AUXX = 2, TF_AUXY = 4, TF_AUXZ = 8, TF_MUNI = 16, TF_PMLZ = 32, TF_MARR = 64, TF_DONE = 128, TF_PML = 256,
              TF_GROUP_E = 512, TF_LAST = 1024;

template <class T>
struct TmaParams {
  // tensor maps, shape index 0 (32x32 tile); shapes 1 (64x16) and 2 (128x8) follow at +1, +2
  const CUtensorMap* mapA[3];   // curl operands, halo box (tw+4) x (th+1)
  const CUtensorMap* mapF[3];   // updated fields, owner box tw x th
  const CUtensorMap* mapM[3];   // per-voxel constitutive arrays (MARR == 1)
  const CUtensorMap* mapW[3];   // W[d] on the slab of axis d
  const CUtensorMap* mapU[3];   // U[d] on the slab of axis next(d)
  T* F[3];
  T* W[3];
  T* U[3];
  long long plane, mplane;
  int px, mpx;
  int n[3];
  int cxp, cy, cz;
  Slab slab[3];
  T dt, idl[3], m_inv;
  const T* sg[3];
  const T* om[3];
  const T* ip[3];
  const WorkItem* items;
  int nitems;
  int nstages;
  int marr;                // the group has per-voxel constitutive arrays
  unsigned int* counter;   // [0] next item, [1] CTAs done (the last one resets both for the next launch)
};

// fused H + E step (step_tma_kernel): the tiles of both half-steps in one persistent launch
struct TmaFuse {
  unsigned long long* done_h;        // per z chunk: consumer warps that have finished an H tile of the chunk (monotonic)
  const unsigned int* cnt_h;         // per z chunk: H tiles of the chunk in this table
  unsigned long long epoch;          // launches so far + 1: the target is epoch * 8 * cnt_h[c]
  int nchunk;
  int* err_flag;
};

// ---- PTX wrappers -------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  while (!mbar_try_wait(bar, parity)) {}
}
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* map, int c0, int c1, int c2, uint32_t bar) {
  asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(dst),
               "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(bar)
               : "memory");
}

template <class T>
__device__ __forceinline__ V4<T> lds4(const unsigned char* base, int off_bytes) {
  return *reinterpret_cast<const V4<T>*>(base + off_bytes);
}

// ----------------------------------------------------------------------------
// Shared-memory geometry (bytes): barriers | 2 ring entries | S stages
// ----------------------------------------------------------------------------
constexpr int TMA_STAGE = tma_stage_bytes(0);
constexpr int TMA_OFF_A = TMA_HDR, TMA_OFF_F = TMA_OFF_A + 3 * TMA_ABOX, TMA_OFF_X = TMA_OFF_F + 3 * TMA_OBOX;   // aux: Wx Uz | Wy Ux | Wz Uy
constexpr int TMA_OFF_M = TMA_OFF_X;     // staged material boxes of an interior tile alias its unused aux slots

struct TmaProducer {     // producer-warp pipeline state
  int s, entry;
  uint32_t eph, fph;
};

// per-tile state of a consumer thread
template <class T>
struct TmaTile {
  int lane_x, row, gx, iy;
  bool act, hasx, hasy;
  bool valid[4];
  Co<T> cx[4], cyc;
  int offA, offAy, offO, offXs;   // byte offsets inside a halo box / owner box / x-slab box
  T zc_s, zc_om, zc_ip;           // z coefficients of plane z0 + lane (handed out by shuffle)
  int fo, mo, xs_off, ys_off;
  V4<T> ax_c, ay_c;               // the z-neighbour plane carried in registers
};

// ----------------------------------------------------------------------------
// producer: all planes of one work item.   GROUP 0: H from curl E, marches DOWN in z; GROUP 1: E from curl H,
// marches UP.  Material factor per tile: scalar m^-1 of the launch, the tile's constant (planner-detected), or
// the per-voxel arrays staged like everything else (interior tiles only: the three boxes use the aux slots an
// interior tile leaves empty; PML tiles with non-constant material stay on step_kernel — a handful per scene).
// ----------------------------------------------------------------------------
template <class T, int GROUP>
__device__ __forceinline__ void tma_produce_item(const TmaParams<T>& p, const WorkItem& it, int lane, int S, uint32_t bars,
                                                 unsigned char* entries, unsigned char* stages, TmaProducer& ps) {
  constexpr int DZ = (GROUP == 0) ? -1 : 1;
  const int si = it.lx_log2 - 3;
  const int tw = 32 << si, th = 32 >> si;
  const int axm = (it.flags >> 4) & 7;     // axes with a PML inside the tile (planner)
  const int zfirst = (DZ > 0) ? it.z0 : it.z0 + it.zn - 1;
  const unsigned zmask = (unsigned)it.zmask;
  const int mmode = p.marr ? ((it.flags & 2) ? TF_MUNI : TF_MARR) : 0;   // 0: the launch's scalar
  const int fx = it.x0 + XO, mx = it.x0 - 1;
  const int ax = (GROUP == 0) ? fx : fx - 4, ay = (GROUP == 0) ? it.y0 : it.y0 - 1;
  const int cx0 = p.slab[0].idx(it.x0), cy0 = p.slab[1].idx(it.y0);
  const uint32_t abytes = (uint32_t)((tw + 4) * (th + 1) * 4), obytes = (uint32_t)(tw * th * 4);
  for (int q = 0; q < it.zn; ++q) {
    const int iz = zfirst + DZ * q;
    const bool first = q == 0;
    const int s = ps.s;
    const uint32_t full = bars + 8 * s, empty = bars + 64 + 8 * s;
    if (lane == 0) {
      mbar_wait(empty, (ps.eph >> s) & 1u);
      if (first) mbar_wait(bars + 128 + 8 * ps.entry, (ps.fph >> ps.entry) & 1u);
    }
    __syncwarp();
    ps.eph ^= 1u << s;
    unsigned char* st = stages + (size_t)s * TMA_STAGE;
    const uint32_t st32 = smem_u32(st);
    const bool auxz = (zmask >> (iz - it.z0)) & 1u;
    // ---- TMA loads, one per lane ----
    uint32_t bytes = 3 * abytes + 3 * obytes;
    if (lane < 3) tma_load_3d(st32 + TMA_OFF_A + lane * TMA_ABOX, p.mapA[lane] + si, ax, ay, iz, full);
    else if (lane < 6) tma_load_3d(st32 + TMA_OFF_F + (lane - 3) * TMA_OBOX, p.mapF[lane - 3] + si, fx, it.y0, iz, full);
    if (axm & 1) {   // x slab: W[0], U[2]
      if (lane == 6) tma_load_3d(st32 + TMA_OFF_X, p.mapW[0] + si, cx0, it.y0 - 1, iz - 1, full);
      if (lane == 7) tma_load_3d(st32 + TMA_OFF_X + TMA_OBOX, p.mapU[2] + si, cx0, it.y0 - 1, iz - 1, full);
      bytes += 2 * obytes;
    }
    if (axm & 2) {   // y slab: W[1], U[0]
      if (lane == 8) tma_load_3d(st32 + TMA_OFF_X + 2 * TMA_OBOX, p.mapW[1] + si, mx, cy0, iz - 1, full);
      if (lane == 9) tma_load_3d(st32 + TMA_OFF_X + 3 * TMA_OBOX, p.mapU[0] + si, mx, cy0, iz - 1, full);
      bytes += 2 * obytes;
    }
    if (auxz) {      // z slab: W[2], U[1]
      const int cz0 = p.slab[2].idx(iz);
      if (lane == 10) tma_load_3d(st32 + TMA_OFF_X + 4 * TMA_OBOX, p.mapW[2] + si, mx, it.y0 - 1, cz0, full);
      if (lane == 11) tma_load_3d(st32 + TMA_OFF_X + 5 * TMA_OBOX, p.mapU[1] + si, mx, it.y0 - 1, cz0, full);
      bytes += 2 * obytes;
    }
    if (mmode == TF_MARR) {   // planner guarantee: axm == 0 here
      if (lane >= 12 && lane < 15) tma_load_3d(st32 + TMA_OFF_M + (lane - 12) * TMA_OBOX, p.mapM[lane - 12] + si, mx, it.y0 - 1, iz - 1, full);
      bytes += 3 * obytes;
    }
    if (first) {
      // neighbour plane of Ax, Ay (carried in registers by the consumers afterwards)
      const uint32_t en32 = smem_u32(entries + (size_t)ps.entry * TMA_ENTRY);
      if (lane == 15) tma_load_3d(en32, p.mapA[0] + si, ax, ay, iz - DZ, full);
      if (lane == 16) tma_load_3d(en32 + TMA_ABOX, p.mapA[1] + si, ax, ay, iz - DZ, full);
      bytes += 2 * abytes;
    }
    if (lane == 0) {
      TmaHdr h;
      h.x0 = it.x0; h.y0 = it.y0; h.xw = it.xw; h.yh = it.yh;
      h.iz = iz; h.lx_log2 = it.lx_log2; h.entry = ps.entry;
      h.flags = (first ? TF_FIRST : 0) | ((axm & 1) ? TF_AUXX : 0) | ((axm & 2) ? TF_AUXY : 0) | ((axm & 4) ? TF_PMLZ : 0) |
                (auxz ? TF_AUXZ : 0) | mmode | (axm ? TF_PML : 0) | (GROUP == 1 ? TF_GROUP_E : 0) | (q == it.zn - 1 ? TF_LAST : 0);
      h.q = q; h.z0 = it.z0; h.zn = it.zn; h.cx0 = cx0;
      h.mu0 = (float)it.mu[0]; h.mu1 = (float)it.mu[1]; h.mu2 = (float)it.mu[2];
      h.chunk = it.chunk;
      *reinterpret_cast<TmaHdr*>(st) = h;
      mbar_arrive_expect_tx(full, bytes);     // release: the header store precedes it
    }
    if (first) { ps.fph ^= 1u << ps.entry; ps.entry ^= 1; }
    ps.s = (s + 1 == S) ? 0 : s + 1;
  }
}

// consumer: first plane of a tile — thread mapping, PML coefficients, the carried neighbour plane
template <class T, int GROUP, bool RAGGED>
__device__ __forceinline__ void tma_consume_first(const TmaParams<T>& p, const TmaHdr& h, TmaTile<T>& t, const unsigned char* entries,
                                                  uint32_t bars, int tid) {
  constexpr int IC = (GROUP == 0) ? 1 : -1;
  const int lxl = h.lx_log2, LX = 1 << lxl;
  const int tw = 4 << lxl;
  t.lane_x = tid & (LX - 1);
  t.row = tid >> lxl;
  t.gx = h.x0 + 4 * t.lane_x;
  t.iy = h.y0 + t.row;
  const int lanes = (h.xw + 3) >> 2;
  t.act = (t.lane_x < lanes) && (t.row < h.yh);
  if (RAGGED) {
#pragma unroll
    for (int e = 0; e < 4; ++e) t.valid[e] = (4 * t.lane_x + e) < h.xw;
  }
  const int pitch = (tw + 4) * 4;
  t.offA = (GROUP == 0) ? (t.row * pitch + 16 * t.lane_x) : ((t.row + 1) * pitch + 16 + 16 * t.lane_x);
  t.offAy = t.offA + IC * pitch;
  t.offO = (t.row * tw + 4 * t.lane_x) * 4;
  t.fo = (t.gx + XO) + p.px * t.iy;
  t.mo = (t.gx - 1) + p.mpx * (t.iy - 1);
  // PML coefficients of the thread's 4 columns and of its row: global loads, once per tile (the producer keeps
  // the stages filling meanwhile); z coefficients of the tile's planes: lane j of every warp holds plane j,
  // handed out by shuffle in tma_consume_plane
#pragma unroll
  for (int e = 0; e < 4; ++e) { t.cx[e].s = T(0); t.cx[e].om = T(1); t.cx[e].ip = T(1); }
  t.cyc.s = T(0); t.cyc.om = T(1); t.cyc.ip = T(1);
  t.hasx = false;
  if (t.act) {
    if (h.flags & TF_AUXX) {
      const V4<T> sv = ld4(p.sg[0] + t.gx - 1), ov = ld4(p.om[0] + t.gx - 1), iv = ld4(p.ip[0] + t.gx - 1);
#pragma unroll
      for (int e = 0; e < 4; ++e) { t.cx[e].s = sv.v[e]; t.cx[e].om = ov.v[e]; t.cx[e].ip = iv.v[e]; t.hasx |= (sv.v[e] != T(0)); }
    }
    if (h.flags & TF_AUXY) { t.cyc.s = p.sg[1][t.iy - 1]; t.cyc.om = p.om[1][t.iy - 1]; t.cyc.ip = p.ip[1][t.iy - 1]; }
  }
  t.zc_s = T(0); t.zc_om = T(1); t.zc_ip = T(1);
  if (h.flags & TF_PMLZ) {
    const int zq = h.z0 + min((int)(tid & 31), h.zn - 1);
    t.zc_s = p.sg[2][zq - 1]; t.zc_om = p.om[2][zq - 1]; t.zc_ip = p.ip[2][zq - 1];
  }
  const unsigned char* en = entries + (size_t)h.entry * TMA_ENTRY;
  if (t.act) {
    t.ax_c = lds4<T>(en, t.offA);
    t.ay_c = lds4<T>(en + TMA_ABOX, t.offA);
  }
  t.hasy = t.act && t.cyc.s != T(0);
  if (t.hasx) {
    // the x slab is compact: W / U of column x sit at idx(x), and the staged box starts at idx(x0)
    const int ci = p.slab[0].idx(t.gx);
    t.xs_off = ci + p.cxp * (t.iy - 1);
    t.offXs = (t.row * tw + (ci - h.cx0)) * 4;
  }
  if (t.hasy) t.ys_off = (t.gx - 1) + p.mpx * p.slab[1].idx(t.iy);
  __syncwarp();
  if ((tid & 31) == 0) mbar_arrive(bars + 128 + 8 * h.entry);
}

// consumer: one staged plane
template <class T, int GROUP, bool RAGGED>
__device__ __forceinline__ void tma_consume_plane(const TmaParams<T>& p, const TmaHdr& h, TmaTile<T>& t, const unsigned char* st) {
  const T dt = p.dt, idx_ = p.idl[0], idy_ = p.idl[1], idz_ = p.idl[2];
  // z coefficients of this plane (warp-uniform)
  Co<T> czc;
  {
    const int src = h.iz - h.z0;
    czc.s = __shfl_sync(0xffffffffu, t.zc_s, src);
    czc.om = __shfl_sync(0xffffffffu, t.zc_om, src);
    czc.ip = __shfl_sync(0xffffffffu, t.zc_ip, src);
  }
  const int iz = h.iz;
  if (!t.act) return;
  const bool hasx = t.hasx, hasy = t.hasy, hasz = czc.s != T(0);
  const int offA = t.offA, offAy = t.offAy, offO = t.offO;
  // ---- staged loads (LDS) ----
  const V4<T> ax0 = lds4<T>(st + TMA_OFF_A, offA), ay0 = lds4<T>(st + TMA_OFF_A + TMA_ABOX, offA), az0 = lds4<T>(st + TMA_OFF_A + 2 * TMA_ABOX, offA);
  const V4<T> ax_y = lds4<T>(st + TMA_OFF_A, offAy), az_y = lds4<T>(st + TMA_OFF_A + 2 * TMA_ABOX, offAy);
  const T ay_x = *reinterpret_cast<const T*>(st + TMA_OFF_A + TMA_ABOX + offA + (GROUP == 0 ? 16 : -4));
  const T az_x = *reinterpret_cast<const T*>(st + TMA_OFF_A + 2 * TMA_ABOX + offA + (GROUP == 0 ? 16 : -4));
  V4<T> fx = lds4<T>(st + TMA_OFF_F, offO), fy = lds4<T>(st + TMA_OFF_F + TMA_OBOX, offO), fz = lds4<T>(st + TMA_OFF_F + 2 * TMA_OBOX, offO);
  V4<T> wx = zero4<T>(), uz = zero4<T>(), wy = zero4<T>(), ux = zero4<T>(), wz = zero4<T>(), uy = zero4<T>();
  if (hasx) { wx = lds4<T>(st + TMA_OFF_X, t.offXs); uz = lds4<T>(st + TMA_OFF_X + TMA_OBOX, t.offXs); }
  if (hasy) { wy = lds4<T>(st + TMA_OFF_X + 2 * TMA_OBOX, offO); ux = lds4<T>(st + TMA_OFF_X + 3 * TMA_OBOX, offO); }
  if (hasz) { wz = lds4<T>(st + TMA_OFF_X + 4 * TMA_OBOX, offO); uy = lds4<T>(st + TMA_OFF_X + 5 * TMA_OBOX, offO); }
  T m0[4], m1[4], m2[4];
  if (h.flags & TF_MARR) {
    const V4<T> a0 = lds4<T>(st + TMA_OFF_M, offO), a1 = lds4<T>(st + TMA_OFF_M + TMA_OBOX, offO), a2 = lds4<T>(st + TMA_OFF_M + 2 * TMA_OBOX, offO);
#pragma unroll
    for (int e = 0; e < 4; ++e) { m0[e] = a0.v[e]; m1[e] = a1.v[e]; m2[e] = a2.v[e]; }
  } else {
    const bool u = (h.flags & TF_MUNI) != 0;
#pragma unroll
    for (int e = 0; e < 4; ++e) { m0[e] = u ? h.mu0 : p.m_inv; m1[e] = u ? h.mu1 : p.m_inv; m2[e] = u ? h.mu2 : p.m_inv; }
  }
  // ---- curl, same operation order as step_kernel / the reference (Helpers.jl:286-298) ----
  const V4<T> ax_z = t.ax_c, ay_z = t.ay_c;
  T kx[4], ky[4], kz[4];
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
    const T k0 = dt * (idz_ * (ay_z.v[e] - ay0.v[e]) - idy_ * (az_y.v[e] - az0.v[e]));
    const T k1 = dt * (idx_ * (azx - az0.v[e]) - idz_ * (ax_z.v[e] - ax0.v[e]));
    const T k2 = dt * (idy_ * (ax_y.v[e] - ax0.v[e]) - idx_ * (ayx - ay0.v[e]));
    kx[e] = m0[e] * k0;
    ky[e] = m1[e] * k1;
    kz[e] = m2[e] * k2;
  }
  const long long base = p.plane * (long long)iz + t.fo;
  if (!(h.flags & TF_PML)) {
    // interior tile: F += m^-1 K, exactly step_kernel's MODE 0 arithmetic
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      if (t.valid[e]) { fx.v[e] = fx.v[e] + kx[e]; fy.v[e] = fy.v[e] + ky[e]; fz.v[e] = fz.v[e] + kz[e]; }
    }
    st4(p.F[0] + base, fx);
    st4(p.F[1] + base, fy);
    st4(p.F[2] + base, fz);
  } else {
    Co<T> cyv4[4], czv4[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) { cyv4[e] = t.cyc; czv4[e] = czc; }
    const T zero[4] = {T(0), T(0), T(0), T(0)};
    V4<T> cdummy = zero4<T>();
    cascade<T, false>(kx, cyv4, czv4, t.cx, hasy, hasx, ux, wx, fx, zero, zero, false, zero, cdummy, t.valid);
    cascade<T, false>(ky, czv4, t.cx, cyv4, hasz, hasy, uy, wy, fy, zero, zero, false, zero, cdummy, t.valid);
    cascade<T, false>(kz, t.cx, cyv4, czv4, hasx, hasz, uz, wz, fz, zero, zero, false, zero, cdummy, t.valid);
    st4(p.F[0] + base, fx);
    st4(p.F[1] + base, fy);
    st4(p.F[2] + base, fz);
    if (hasx) {
      const long long xsl = (long long)p.cxp * p.n[1] * (long long)(iz - 1) + t.xs_off;
      st4(p.W[0] + xsl, wx); st4(p.U[2] + xsl, uz);
    }
    if (hasy) {
      const long long ysl = (long long)p.mpx * p.cy * (long long)(iz - 1) + t.ys_off;
      st4(p.W[1] + ysl, wy); st4(p.U[0] + ysl, ux);
    }
    if (hasz) {
      const long long zsl = p.mplane * (long long)p.slab[2].idx(iz) + t.mo;
      st4(p.W[2] + zsl, wz); st4(p.U[1] + zsl, uy);
    }
  }
  t.ax_c = ax0;
  t.ay_c = ay0;
}

__device__ __forceinline__ void tma_init_barriers(uint32_t bars, int S, int tid) {
  // barriers: full[S] @0, empty[S] @64, efree[2] @128
  if (tid == 0) {
    for (int s = 0; s < S; ++s) { mbar_init(bars + 8 * s, 1); mbar_init(bars + 64 + 8 * s, TMA_CONSUMERS / 32); }
    mbar_init(bars + 128, TMA_CONSUMERS / 32);
    mbar_init(bars + 136, TMA_CONSUMERS / 32);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
}
template <class T>
__device__ __forceinline__ void tma_init_tile(TmaTile<T>& t) {
  t.lane_x = t.row = t.gx = t.iy = 0;
  t.act = t.hasx = t.hasy = false;
#pragma unroll
  for (int e = 0; e < 4; ++e) { t.valid[e] = true; t.cx[e].s = T(0); t.cx[e].om = T(1); t.cx[e].ip = T(1); }
  t.cyc.s = T(0); t.cyc.om = T(1); t.cyc.ip = T(1);
  t.offA = t.offAy = t.offO = t.offXs = 0;
  t.zc_s = T(0); t.zc_om = T(1); t.zc_ip = T(1);
  t.fo = t.mo = t.xs_off = t.ys_off = 0;
  t.ax_c = zero4<T>(); t.ay_c = zero4<T>();
}
// the producer's goodbye: a sentinel stage for the consumers, and the last CTA re-arms the work counter
__device__ __forceinline__ void tma_producer_finish(uint32_t bars, unsigned char* stages, TmaProducer& ps, unsigned int* counter) {
  mbar_wait(bars + 64 + 8 * ps.s, (ps.eph >> ps.s) & 1u);
  TmaHdr h = {};
  h.flags = TF_DONE;
  *reinterpret_cast<TmaHdr*>(stages + (size_t)ps.s * TMA_STAGE) = h;
  mbar_arrive(bars + 8 * ps.s);
  __threadfence();
  const unsigned done = atomicAdd(counter + 1, 1u);
  if (done == gridDim.x - 1) { counter[0] = 0; counter[1] = 0; __threadfence(); }
}

// ----------------------------------------------------------------------------
// One half-step: interior + PML tiles of one field group.   RAGGED: Nx % 4 != 0
// ----------------------------------------------------------------------------
template <class T, int GROUP, bool RAGGED>
__global__ void KHR_TMA_BOUNDS pml_tma_kernel(const __grid_constant__ TmaParams<T> p) {
  static_assert(sizeof(T) == 4, "Float32 only");
  extern __shared__ __align__(1024) unsigned char smem[];
  const int S = p.nstages;
  const uint32_t bars = smem_u32(smem);
  unsigned char* entries = smem + 1024;
  unsigned char* stages = entries + 2 * TMA_ENTRY;
  const int tid = threadIdx.x;
  tma_init_barriers(bars, S, tid);

  if (tid >= TMA_CONSUMERS) {
    // ============================ producer warp ============================
    // Nothing on this warp's path waits for global memory: the next work item is claimed and read one
    // item ahead (the atomic and the loads are in flight while the current item's planes are issued),
    // and everything a plane needs beyond the item comes from registers.
    const int lane = tid - TMA_CONSUMERS;
    TmaProducer ps;
    ps.s = 0; ps.entry = 0; ps.fph = 3; ps.eph = 0;   // parities start "free": waiting for parity 1 passes at once
    for (int q = 0; q < S; ++q) ps.eph |= 1u << q;
    unsigned idx = 0, idx_next = 0;
    if (lane == 0) { idx = atomicAdd(p.counter, 1u); idx_next = atomicAdd(p.counter, 1u); }
    idx = __shfl_sync(0xffffffffu, idx, 0);
    idx_next = __shfl_sync(0xffffffffu, idx_next, 0);
    WorkItem it, it_next;
    if (idx < (unsigned)p.nitems) it = p.items[idx];
    while (idx < (unsigned)p.nitems) {
      if (idx_next < (unsigned)p.nitems) it_next = p.items[idx_next];
      unsigned idx_next2 = 0;
      if (lane == 0) idx_next2 = atomicAdd(p.counter, 1u);
      tma_produce_item<T, GROUP>(p, it, lane, S, bars, entries, stages, ps);
      idx = idx_next; it = it_next;
      idx_next = __shfl_sync(0xffffffffu, idx_next2, 0);
    }
    if (lane == 0) tma_producer_finish(bars, stages, ps, p.counter);
    return;
  }

  // ============================== consumers ==============================
  int s = 0;
  uint32_t fphase = 0;        // parity of full[s], tracked per stage in bits
  TmaTile<T> t;
  tma_init_tile(t);
  for (;;) {
    mbar_wait(bars + 8 * s, (fphase >> s) & 1u);
    fphase ^= 1u << s;
    const unsigned char* st = stages + (size_t)s * TMA_STAGE;
    const TmaHdr h = *reinterpret_cast<const TmaHdr*>(st);
    if (h.flags & TF_DONE) break;
    if (h.flags & TF_FIRST) tma_consume_first<T, GROUP, RAGGED>(p, h, t, entries, bars, tid);
    tma_consume_plane<T, GROUP, RAGGED>(p, h, t, st);
    __syncwarp();
    if ((tid & 31) == 0) mbar_arrive(bars + 64 + 8 * s);
    s = (s + 1 == S) ? 0 : s + 1;
  }
}

// ----------------------------------------------------------------------------
// One whole time step: the interior + PML tiles of BOTH half-steps in one persistent launch.  The table
// holds them in z-chunk order with the E tiles one chunk behind the H tiles — H(0) H(1) E(0) H(2) E(1) ... —
// so that E(c) finds the H planes of chunk c (written one chunk ago) and its own E planes (read by H(c)) in
// the 126 MB L2 instead of HBM, and no grid-wide drain separates the half-steps.  Ordering: an E tile of
// chunk c reads H of chunks c-1, c and overwrites E planes that H(c-1), H(c) read; its producer therefore
// waits until every H tile of those two chunks has finished (per-chunk completion counters bumped by the
// consumer warps after their last store, acquire-polled by the producer only — consumers never spin), then
// issues its TMA loads.  Dependencies always point to tiles claimed earlier, and H tiles never wait, so the
// persistent grid cannot deadlock.  H tiles need E of the previous step only: the previous launch.
// ----------------------------------------------------------------------------
template <class T, bool RAGGED>
__global__ void KHR_TMA_BOUNDS step_tma_kernel(const __grid_constant__ TmaParams<T> ph, const __grid_constant__ TmaParams<T> pe,
                                                                   const __grid_constant__ TmaFuse fu) {
  static_assert(sizeof(T) == 4, "Float32 only");
  extern __shared__ __align__(1024) unsigned char smem[];
  const int S = ph.nstages;
  const uint32_t bars = smem_u32(smem);
  unsigned char* entries = smem + 1024;
  unsigned char* stages = entries + 2 * TMA_ENTRY;
  const int tid = threadIdx.x;
  tma_init_barriers(bars, S, tid);

  if (tid >= TMA_CONSUMERS) {
    const int lane = tid - TMA_CONSUMERS;
    TmaProducer ps;
    ps.s = 0; ps.entry = 0; ps.fph = 3; ps.eph = 0;
    for (int q = 0; q < S; ++q) ps.eph |= 1u << q;
    unsigned idx = 0, idx_next = 0;
    if (lane == 0) { idx = atomicAdd(ph.counter, 1u); idx_next = atomicAdd(ph.counter, 1u); }
    idx = __shfl_sync(0xffffffffu, idx, 0);
    idx_next = __shfl_sync(0xffffffffu, idx_next, 0);
    WorkItem it, it_next;
    if (idx < (unsigned)ph.nitems) it = ph.items[idx];
    while (idx < (unsigned)ph.nitems) {
      if (idx_next < (unsigned)ph.nitems) it_next = ph.items[idx_next];
      unsigned idx_next2 = 0;
      if (lane == 0) idx_next2 = atomicAdd(ph.counter, 1u);
      if ((it.flags >> 8) & 1) {
        // E tile: wait for the H tiles of chunks c-1 and c (two lanes poll the two counters at once)
        if (lane < 2) {
          const int c = it.chunk - 1 + lane;
          if (c >= 0 && c < fu.nchunk) {
            const unsigned long long target = fu.epoch * 8ull * (unsigned long long)fu.cnt_h[c];
            unsigned spins = 0;
            while (ld_acquire_u64(fu.done_h + c) < target) {
              __nanosleep(64);
              if (++spins > (1u << 22)) { *fu.err_flag = 2; break; }   // never hang the GPU (~1 s)
            }
          }
        }
        __syncwarp();
        asm volatile("fence.proxy.async;" ::: "memory");   // the H stores of other SMs (generic proxy) before this tile's TMA reads
        tma_produce_item<T, 1>(pe, it, lane, S, bars, entries, stages, ps);
      } else {
        tma_produce_item<T, 0>(ph, it, lane, S, bars, entries, stages, ps);
      }
      idx = idx_next; it = it_next;
      idx_next = __shfl_sync(0xffffffffu, idx_next2, 0);
    }
    if (lane == 0) tma_producer_finish(bars, stages, ps, ph.counter);
    return;
  }

  int s = 0;
  uint32_t fphase = 0;
  TmaTile<T> t;
  tma_init_tile(t);
  for (;;) {
    mbar_wait(bars + 8 * s, (fphase >> s) & 1u);
    fphase ^= 1u << s;
    const unsigned char* st = stages + (size_t)s * TMA_STAGE;
    const TmaHdr h = *reinterpret_cast<const TmaHdr*>(st);
    if (h.flags & TF_DONE) break;
    if (h.flags & TF_GROUP_E) {
      if (h.flags & TF_FIRST) tma_consume_first<T, 1, RAGGED>(pe, h, t, entries, bars, tid);
      tma_consume_plane<T, 1, RAGGED>(pe, h, t, st);
    } else {
      if (h.flags & TF_FIRST) tma_consume_first<T, 0, RAGGED>(ph, h, t, entries, bars, tid);
      tma_consume_plane<T, 0, RAGGED>(ph, h, t, st);
    }
    __syncwarp();
    if ((tid & 31) == 0) {
      mbar_arrive(bars + 64 + 8 * s);
      if ((h.flags & (TF_LAST | TF_GROUP_E)) == TF_LAST) {
        // this warp has stored the last plane of an H tile: publish (release) one completion for the chunk
        __threadfence();
        atomicAdd(fu.done_h + h.chunk, 1ull);
      }
    }
    s = (s + 1 == S) ? 0 : s + 1;
  }
}

}  // namespace khr
