// ============================================================================
// khronos_b200.cu — host side of libkhronos_b200.so: context, storage layout,
// work-table planner, per-step orchestration, DFT, halo exchange, C ABI.
// Reference call sites are cited in include/khronos_b200.h next to each entry.
//
// There is deliberately NO CPU fallback in this file: every compute entry
// point launches the sm_100a kernels of step_kernels.cuh or fails with an error.
// ============================================================================
#include "../../include/khronos_b200.h"
#include "step_kernels.cuh"
#include "post_kernels.cuh"
#include "geom_kernels.cuh"
#include "pml_tma.cuh"

#include <dlfcn.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

namespace khr {

static thread_local std::string g_err;
static int32_t fail(const std::string& m) {
  g_err = m;
  return 1;
}
#define CUDA_OK(expr)                                                                               \
  do {                                                                                              \
    cudaError_t _e = (expr);                                                                        \
    if (_e != cudaSuccess)                                                                          \
      throw std::string(#expr) + ": " + cudaGetErrorString(_e) + " (" + __FILE__ + ":" + std::to_string(__LINE__) + ")"; \
  } while (0)

static inline int round_up(int a, int b) { return (a + b - 1) / b * b; }

// ---- NCCL through dlopen: the library must not pull a second libnccl into a
// process that already carries one (torch bundles its own).
struct Id128 {
  char b[128];
};
struct NcclApi {
  void* h = nullptr;
  int (*GetUniqueId)(void*) = nullptr;
  int (*CommInitRank)(void**, int, /*ncclUniqueId by value, 128 bytes*/ Id128, int) = nullptr;
  int (*CommDestroy)(void*) = nullptr;
  int (*Send)(const void*, size_t, int, int, void*, cudaStream_t) = nullptr;
  int (*Recv)(void*, size_t, int, int, void*, cudaStream_t) = nullptr;
  int (*AllReduce)(const void*, void*, size_t, int, int, void*, cudaStream_t) = nullptr;
  int (*CommSplit)(void*, int, int, void**, void*) = nullptr;
  int (*GroupStart)() = nullptr;
  int (*GroupEnd)() = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
};
static NcclApi g_nccl;
static void load_nccl() {
  if (g_nccl.h) return;
  void* h = nullptr;
  // 1) symbols already in the process (torch's bundled NCCL)
  if (dlsym(RTLD_DEFAULT, "ncclCommInitRank")) h = RTLD_DEFAULT;
  const char* names[] = {"libnccl.so.2", "libnccl.so"};
  for (int i = 0; !h && i < 2; ++i) h = dlopen(names[i], RTLD_NOW | RTLD_GLOBAL);
  if (!h) throw std::string("NCCL not found (dlopen libnccl.so.2 failed): ") + (dlerror() ? dlerror() : "");
  auto sym = [&](const char* n) {
    void* p = dlsym(h, n);
    if (!p) throw std::string("NCCL symbol missing: ") + n;
    return p;
  };
  g_nccl.GetUniqueId = (int (*)(void*))sym("ncclGetUniqueId");
  g_nccl.CommInitRank = (int (*)(void**, int, Id128, int))sym("ncclCommInitRank");
  g_nccl.CommDestroy = (int (*)(void*))sym("ncclCommDestroy");
  g_nccl.Send = (int (*)(const void*, size_t, int, int, void*, cudaStream_t))sym("ncclSend");
  g_nccl.Recv = (int (*)(void*, size_t, int, int, void*, cudaStream_t))sym("ncclRecv");
  g_nccl.AllReduce = (int (*)(const void*, void*, size_t, int, int, void*, cudaStream_t))sym("ncclAllReduce");
  g_nccl.CommSplit = (int (*)(void*, int, int, void**, void*))dlsym(h, "ncclCommSplit");   // NCCL >= 2.18
  g_nccl.GroupStart = (int (*)())sym("ncclGroupStart");
  g_nccl.GroupEnd = (int (*)())sym("ncclGroupEnd");
  g_nccl.GetErrorString = (const char* (*)(int))sym("ncclGetErrorString");
  g_nccl.h = h;
}
#define NCCL_OK(expr)                                                                      \
  do {                                                                                     \
    int _r = (expr);                                                                       \
    if (_r != 0) throw std::string(#expr) + ": " + khr::g_nccl.GetErrorString(_r);              \
  } while (0)

// ----------------------------------------------------------------------------
struct Base {
  virtual ~Base() {}
  khr_grid_desc g;
  virtual void set_pml_sigma(int group, int axis, const void* s, int len) = 0;
  virtual void set_grid_spacing(int axis, const void* d, int len) = 0;
  virtual void geometry_rasterize(const khr_object* objs, int nobj, int kinds_mask, int smoothing, const double* origins18,
                                  int64_t* smoothed3) = 0;
  virtual void material_read(int kind, int comp, void* out) = 0;
  virtual void set_material_scalar(int kind, double v) = 0;
  virtual void set_material_array(int kind, int comp, const void* dense) = 0;
  virtual int pole_register(double omega0, double gamma, const void* sigma) = 0;
  virtual int source_register(int comp, const int32_t* start, const int32_t* dims, const void* amp, int kind,
                              const double* tp) = 0;
  virtual void source_set_amplitude(int id, double re, double im) = 0;
  virtual int monitor_register(int comp, const int32_t* s, const int32_t* e, int nf, const double* f, int dec) = 0;
  virtual void finalize() = 0;
  virtual void step(int n) = 0;
  virtual void step_h() = 0;
  virtual void step_e() = 0;
  virtual void dft_update(int group, double time) = 0;
  virtual void reset_fields() = 0;
  virtual void comm_init(const void* id, int nranks, int rank) = 0;
  virtual void comm_split_from(Base* parent) = 0;
  void* comm_handle = nullptr;
  virtual void halo_exchange(int group) = 0;
  virtual void field_read(int comp, void* out) = 0;
  virtual void field_write(int comp, const void* in) = 0;
  virtual void field_view(int comp, void** p, int64_t* stride, int64_t* off) = 0;
  virtual void monitor_read(int id, void* out) = 0;
  virtual void monitor_view(int id, void** p, int64_t* dims) = 0;
  virtual double monitor_norm(int id) = 0;
  virtual void monitor_norms(double* out, int count) = 0;
  virtual void flux(const int32_t* ids4, int normal_axis, double* out, int nfreq) = 0;
  virtual void near2far(const int32_t* ids4, int normal_axis, double normal_sign, double eps, double mu, const double* base12,
                        const double* freqs, int nfreq, const double* obs, int nobs, double* out) = 0;
  virtual void mode_overlap(const int32_t* ids4, int normal_axis, const double* mode, int n1, int n2, int nfreq, double* out5) = 0;
  virtual void diffraction(const int32_t* ids4, int normal_axis, int max_order, double L1, double L2, double kinc1, double kinc2,
                           const double* freqs, int nfreq, double* power, int32_t* propagating) = 0;
  virtual void sync() = 0;
  virtual void census(int64_t* c) = 0;
  virtual void set_profiling(int on) = 0;
  virtual int kernel_stat(int idx, khr_kernel_stat* out) = 0;
  virtual void comm_stat(double* wait_ms, int64_t* exchanges) = 0;
  virtual void graph_info(int64_t* kernels_per_graph, int64_t* replays) = 0;
  bool periodic[3] = {false, false, false};
  // complex fields (Bloch boundaries): this context holds the real parts, `partner` (owned by the
  // khr_ctx) the imaginary parts of every field; bloch_kl[a] = k * L of the axis
  bool in_pair = false, is_imag = false;
  double bloch_kl[3] = {0.0, 0.0, 0.0};
  virtual void set_partner(Base* p) = 0;
  bool registered_any = false;
  int64_t timestep = 0;
  int sources_mode = -1;
  bool sources_active = true;
  cudaStream_t stream = nullptr, comm_stream = nullptr;
  double last_ms = 0;
  int64_t last_launches = 0;
  int64_t dev_bytes = 0;
  int device = 0;
};

template <class T>
struct TimeSrc {
  int kind;
  T fcen, width, peak, cutoff;
  T host_re = 0, host_im = 0;
};

// Sources/TimeSources.jl:61-64 (CW), :123-132 (Gaussian): every factor multiplies the
// imaginary part left to right in T; exp(i x) = (cos x, sin x) evaluated in T.
template <class T>
static void eval_time_source(const TimeSrc<T>& s, double t, T* re, T* im) {
  const T pi_T = (T)3.141592653589793;
  if (s.kind == KHR_TIME_CW) {
    T a = -T(1);
    a = a * T(2);
    a = a * pi_T;
    a = a * s.fcen;
    a = a * (T)t;
    *re = std::cos(a);
    *im = std::sin(a);
  } else if (s.kind == KHR_TIME_GAUSSIAN) {
    T tt = (T)t - s.peak;
    if (tt > s.cutoff) { *re = 0; *im = 0; return; }
    T two_pi = T(2) * pi_T;
    T env = std::exp((-tt * tt) / (T(2) * s.width * s.width));
    T a = -two_pi;
    a = a * s.fcen;
    a = a * tt;
    *re = env * std::cos(a);
    *im = env * std::sin(a);
  } else {
    if ((T)t > s.cutoff) { *re = 0; *im = 0; return; }
    *re = s.host_re;
    *im = s.host_im;
  }
}

template <class T>
struct Impl : Base {
  // ---- geometry of the storage ----
  int N[3];          // local cells (Nx, Ny, Nz_local)
  int PX, PY, PZ;    // ghosted field box
  int MPX;           // material row pitch
  size_t fsize, msize;
  T dt, dl[3];
  // ---- device arrays ----
  T* F[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
  T* m_arr[2][3] = {{nullptr, nullptr, nullptr}, {nullptr, nullptr, nullptr}};  // [0]=mu_inv (H group) [1]=eps_inv
  T m_scalar[2] = {T(1), T(1)};
  T* sigM[2][3] = {{nullptr, nullptr, nullptr}, {nullptr, nullptr, nullptr}};   // [0]=sigma_B [1]=sigma_D
  T* Cst[2][3] = {{nullptr, nullptr, nullptr}, {nullptr, nullptr, nullptr}};
  T* Dst[3] = {nullptr, nullptr, nullptr};  // D on dispersive / Kerr voxels
  T* idv[3] = {nullptr, nullptr, nullptr};  // non-uniform grid: inv(Δ[i]) per local cell (null: uniform axis)
  bool nonuniform = false;
  T* chi3 = nullptr;                        // Kerr coefficient (Geometry.jl:610-660), material layout
  int chi3_box[6] = {1, 1, 1, 0, 0, 0};
  std::vector<uint8_t> chi3_mask;
  T* W[2][3] = {{nullptr, nullptr, nullptr}, {nullptr, nullptr, nullptr}};
  T* U[2][3] = {{nullptr, nullptr, nullptr}, {nullptr, nullptr, nullptr}};
  size_t slab_elems[3] = {0, 0, 0};
  T* coef[2][3][3] = {};  // [group][axis][sg|om|ip]
  std::vector<T> h_sig[2][3];  // per-cell sigma (value at 2i-1), local cells, host
  bool have_sigma[2][3] = {{false, false, false}, {false, false, false}};
  Slab slab[3];
  int lo_end[3] = {0, 0, 0}, hi_start[3] = {0, 0, 0};
  int cxp = 0, cy = 0, cz = 0;
  int sd_box[2][6];  // bounding boxes (local cells) of non-zero sigma_B / sigma_D, empty if lo > hi
  bool has_sd[2] = {false, false};

  struct Pole {
    T* sigma = nullptr;
    T* P[2][3];  // ping-pong
    int cur = 0;
    T g1i, g1, cp, cd;
    int box[6];
  };
  std::vector<Pole> poles;
  struct Source {
    int comp;
    int s[3], d[3];  // local start (z already shifted), extent
    T* amp = nullptr;
    int* slot = nullptr;
    TimeSrc<T> ts;
    T ao_re = 0, ao_im = 0;
  };
  std::vector<Source> sources;
  std::vector<SrcDesc<T>> h_src_ext[2];
  SrcDesc<T>* d_src_ext[2] = {nullptr, nullptr};
  size_t src_ext_cap[2] = {0, 0};
  std::vector<PoleDesc<T>> h_pole_ext;
  PoleDesc<T>* d_pole_ext = nullptr;
  T* Tsrc[2] = {nullptr, nullptr};  // flux accumulators of the source voxels per group
  size_t nslots[2] = {0, 0};
  struct Monitor {
    int comp;
    int s[3], e[3], n[3];  // global index box
    std::vector<T> freqs;
    T* d_freqs = nullptr;
    T* M = nullptr;
    int decimation;
    size_t elems;  // complex elements
    bool local;    // any part on this rank
  };
  std::vector<Monitor> monitors;
  MonDesc<T>* d_mons = nullptr;
  std::vector<int> mon_local_nz;

  // work tables: [group][mode][lx index]
  struct Table {
    std::vector<WorkItem> items;
    std::vector<double> cost;  // bytes model per item (launch-order heuristic)
    WorkItem* d = nullptr;
    // bytes model and live CUDA-event timing of this table's launches
    int64_t cells = 0;
    int64_t uniform_items = 0;  // tiles whose per-voxel material arrays are constant
    double alg_bytes = 0;      // compulsory bytes of this implementation
    double ref_bytes = 0;      // the reference's byte model
    unsigned int* d_ctr = nullptr;   // work counter of the persistent TMA kernel (pml_tma.cuh)
    std::vector<cudaEvent_t> ev;  // pairs
    size_t ev_used = 0;
    double total_ms = 0;
    int64_t nlaunch = 0;
  };
  bool profiling = false, serial_prof = false;
  std::vector<std::vector<uint8_t>> pole_mask;  // host non-zero masks (bytes model only)
  std::vector<uint8_t> sd_mask[2];
  // phase 0 = boundary planes that feed the halo exchange, phase 1 = the rest
  // table index: 0 interior; 1 / 2 / 4 PML on x / y / z only; 7 PML on several axes;
  // 8 "full" (sources, conductivity, poles); 3, 5, 6 unused (folded into 7)
  // tables MBASE + m hold the tiles of class m (interior / PML) whose per-voxel material arrays are
  // constant over the tile: they run the MARR = 2 kernels (values in the work item, 12 registers
  // fewer, no material loads) next to the remaining tiles of the class
  static constexpr int MBASE = 10, NTAB = 2 * MBASE, NSIDE = 14;   // class 9: MODE 2 tiles outside the PML (AXM = 0 variant)
  bool split_uniform = true;
  // CUDA graph of one time step (single GPU): the kernels of both half-steps on their streams are captured once;
  // per step only the MODE 2 kernel nodes get new parameters (source amplitudes, P^n / P^{n-1} pointers) and the
  // graph is replayed — one host launch per step instead of 6-8 launches + ~10 event operations.  The reference
  // captures CUDA graphs of its step for the same reason (Kernels.jl:99-145).  Bit-identical; on B200 the host is
  // not the limiter of this path (it runs several steps ahead of the device), and the replayed graph loses the
  // stream priorities that put the PML kernels first, so it measures slower and stays opt-in.
  bool graph_on = false, capturing = false;   // measured (profiles/r02_graph_ab.txt): replay is 3-5 % slower than the direct launches -> opt-in, KHR_GRAPH=1
  cudaGraph_t step_graph = nullptr;
  cudaGraphExec_t step_gexec = nullptr;
  struct DynNode { cudaGraphNode_t node; int gq; const WorkItem* items; cudaKernelNodeParams kp; };
  std::vector<DynNode> dyn_nodes;
  int64_t graph_kernels = 0, graph_replays = 0;
  int plain_steps_done = 0;
  bool full_nopml = false;   // KHR_FULL_NOPML=1: measured no gain (a second latency-bound launch), profiles/r02_mode2_ab.txt
  int full_split = 2;      // MODE 2 tiles: rows per CTA = tile rows / full_split, threads = 256 / full_split
  // sweep mode (KHR_SWEEP=1): one grid per time step with the interior + PML tiles of both half-steps
  // in z-chunk-major order (sweep_kernel, step_kernels.cuh); needs the chain mode's counters
  bool sweep = false;
  Table sweep_tab;
  Table tab[2][2][NTAB];
  Table fix_tab[2];   // [group]: tiles outside the PML whose source / pole / Kerr voxels are rewritten by fixup_kernel after the marching kernels
  // off by default (KHR_FIXUP=1 turns it on): bit-identical, but slower on both workloads it was built for
  // (uled 46.0 -> 43.2, waveguide_mode 50.6 -> 48.5 Gcells/s, profiles/r02_mode2_ab.txt)
  bool fixup_on = false, fix_ok[2] = {false, false};
  Table tma_tab[2];   // [group]: interior + PML tiles of phase 1 handled by the persistent TMA kernel (pml_tma.cuh)
  // fused step (KHR_FUSE=1): both half-steps' TMA tiles in one launch, E one z chunk behind H (step_tma_kernel)
  bool tma_fuse = false, skip_tma_launch = false;
  int fuse_lag = 1;
  Table fuse_tab;
  unsigned long long* d_fuse_done = nullptr;
  unsigned int* d_fuse_cnt = nullptr;
  unsigned long long fuse_epoch = 0;
  cudaStream_t side[NSIDE] = {};
  cudaEvent_t ev_fork = nullptr, ev_fix = nullptr, ev_join[NSIDE] = {};
  bool axis_spec = false;  // measured slower on B200 (profiles/r01_axis_spec_pdl_ab.txt): more launches, more tails
  static int side_of(int mm) {
    const int m = mm % MBASE;
    return (m == 1 ? 0 : m == 2 ? 1 : m == 4 ? 2 : m == 7 ? 3 : m == 8 ? 4 : m == 9 ? 5 : 6) + (mm >= MBASE ? 7 : 0);
  }
  bool main_heaviest = false;
  // chain mode: all kernels of a step on one stream, launched with programmatic dependent launch;
  // H <-> E ordering is enforced per z chunk by device counters (step_kernels.cuh), so the tail of
  // one half-step overlaps the head of the next
  bool pdl = false;
  int nchunk = 0;
  unsigned long long* d_cnt[2] = {nullptr, nullptr};
  unsigned long long* d_done[2] = {nullptr, nullptr};
  unsigned long long epochs[2] = {0, 0};
  int* h_err = nullptr;  // mapped pinned flag, set by a kernel whose dependency wait timed out
  int* d_err = nullptr;
  int launch_order = 0;
  bool multi_stream = true;
  bool finalized = false;
  // distributed
  void* comm = nullptr;
  cudaEvent_t ev_boundary = nullptr, ev_comm = nullptr, ev_t0 = nullptr, ev_t1 = nullptr;
  int64_t launches = 0;

  Impl<T>* im = nullptr;   // imaginary parts (complex fields); null for real fields and for the imaginary context itself
  cudaEvent_t ev_pair_a = nullptr, ev_pair_b = nullptr;
  void set_partner(Base* p) override {
    im = dynamic_cast<Impl<T>*>(p);
    if (!im) throw std::string("complex fields: partner context has another dtype");
    in_pair = true;
    im->in_pair = true; im->is_imag = true;
    CUDA_OK(cudaEventCreateWithFlags(&ev_pair_a, cudaEventDisableTiming));
    CUDA_OK(cudaEventCreateWithFlags(&ev_pair_b, cudaEventDisableTiming));
  }
  std::vector<void*> allocs;
  T* dalloc(size_t n, bool zero = true) {
    void* p = nullptr;
    size_t bytes = (n + 64) * sizeof(T);
    CUDA_OK(cudaMalloc(&p, bytes));
    if (zero) CUDA_OK(cudaMemsetAsync(p, 0, bytes, stream));
    allocs.push_back(p);
    dev_bytes += (int64_t)bytes;
    return (T*)p;
  }

  Impl(int dev, const khr_grid_desc& gd) {
    g = gd;
    device = dev;
    CUDA_OK(cudaSetDevice(dev));
    CUDA_OK(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking));
    CUDA_OK(cudaStreamCreateWithFlags(&comm_stream, cudaStreamNonBlocking));
    CUDA_OK(cudaEventCreateWithFlags(&ev_boundary, cudaEventDisableTiming));
    CUDA_OK(cudaEventCreateWithFlags(&ev_comm, cudaEventDisableTiming));
    CUDA_OK(cudaEventCreate(&ev_t0));
    CUDA_OK(cudaEventCreate(&ev_t1));
    int prio_lo = 0, prio_hi = 0;
    CUDA_OK(cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi));
    bool use_prio = true;
    if (const char* e = getenv("KHR_STREAM_PRIO")) use_prio = atoi(e) != 0;
    for (int q = 0; q < NSIDE; ++q) {
      // the PML / full kernels are the critical path of a half-step: their CTAs get the SM slots
      // first, the interior kernel (main stream, default priority) fills what is left
      CUDA_OK(cudaStreamCreateWithPriority(&side[q], cudaStreamNonBlocking, (use_prio && (q % 7) < 6) ? prio_hi : prio_lo));
      CUDA_OK(cudaEventCreateWithFlags(&ev_join[q], cudaEventDisableTiming));
    }
    if (const char* e = getenv("KHR_AXIS_SPEC")) axis_spec = atoi(e) != 0;
    if (const char* e = getenv("KHR_MAIN_HEAVIEST")) main_heaviest = atoi(e) != 0;
    if (const char* e = getenv("KHR_ORDER")) launch_order = atoi(e);
    if (const char* e = getenv("KHR_CHAIN")) pdl = atoi(e) != 0;
    if (const char* e = getenv("KHR_SPLIT_UNIFORM")) split_uniform = atoi(e) != 0;
    if (const char* e = getenv("KHR_SWEEP")) sweep = atoi(e) != 0;
    if (const char* e = getenv("KHR_TMA")) tma_policy = atoi(e) != 0 ? 1 : 0;
    if (const char* e = getenv("KHR_TMA_STAGES")) tma_stages_req = atoi(e);
    if (const char* e = getenv("KHR_FUSE")) tma_fuse = atoi(e) != 0;
    if (const char* e = getenv("KHR_GRAPH")) graph_on = atoi(e) != 0;
    if (const char* e = getenv("KHR_FIXUP")) fixup_on = atoi(e) != 0;
    if (const char* e = getenv("KHR_FUSE_LAG")) fuse_lag = std::max(0, atoi(e));
    if (sweep) { pdl = true; multi_stream = false; }
    if (pdl) multi_stream = false;
    CUDA_OK(cudaHostAlloc((void**)&h_err, sizeof(int), cudaHostAllocMapped));
    *h_err = 0;
    CUDA_OK(cudaHostGetDevicePointer((void**)&d_err, h_err, 0));
    CUDA_OK(cudaEventCreateWithFlags(&ev_fork, cudaEventDisableTiming));
    CUDA_OK(cudaEventCreateWithFlags(&ev_fix, cudaEventDisableTiming));
    if (const char* e = getenv("KHR_MULTI_STREAM")) multi_stream = atoi(e) != 0;
    N[0] = gd.n[0]; N[1] = gd.n[1]; N[2] = gd.nz_local;
    PX = round_up(N[0] + 36, 32);
    PY = N[1] + 2;
    PZ = N[2] + 2;
    MPX = round_up(N[0], 32);
    fsize = (size_t)PX * PY * PZ;
    msize = (size_t)MPX * N[1] * N[2];
    dt = (T)gd.dt;
    for (int a = 0; a < 3; ++a) dl[a] = (T)gd.dl[a];
    for (int c = 0; c < 6; ++c) F[c] = dalloc(fsize);
    for (int gq = 0; gq < 2; ++gq) { sd_box[gq][0] = 1; sd_box[gq][3] = 0; }
  }
  ~Impl() override {
    cudaSetDevice(device);
    cudaDeviceSynchronize();
    for (void* p : allocs) cudaFree(p);
    if (h_norms) cudaFreeHost(h_norms);
    if (h_norm_nb) cudaFreeHost(h_norm_nb);
    if (h_err) cudaFreeHost(h_err);
    if (comm && g_nccl.CommDestroy) g_nccl.CommDestroy(comm);
    cudaEventDestroy(ev_boundary); cudaEventDestroy(ev_comm); cudaEventDestroy(ev_t0); cudaEventDestroy(ev_t1);
    if (step_gexec) cudaGraphExecDestroy(step_gexec);
    if (step_graph) cudaGraphDestroy(step_graph);
    if (ev_fork) cudaEventDestroy(ev_fork);
    if (ev_fix) cudaEventDestroy(ev_fix);
    if (ev_pair_a) cudaEventDestroy(ev_pair_a);
    if (ev_pair_b) cudaEventDestroy(ev_pair_b);
    for (int q = 0; q < NSIDE; ++q) {
      if (ev_join[q]) cudaEventDestroy(ev_join[q]);
      if (side[q]) cudaStreamDestroy(side[q]);
    }
    for_tables([&](Table& t, int, int, int) { for (cudaEvent_t e : t.ev) cudaEventDestroy(e); t.ev.clear(); });
    for (cudaEvent_t e : sweep_tab.ev) cudaEventDestroy(e);
    for (cudaEvent_t e : halo_ev) cudaEventDestroy(e);
    for (int gq = 0; gq < 2; ++gq) for (cudaEvent_t e : tma_tab[gq].ev) cudaEventDestroy(e);
    for (cudaEvent_t e : fuse_tab.ev) cudaEventDestroy(e);
    for (int gq = 0; gq < 2; ++gq) for (cudaEvent_t e : fix_tab[gq].ev) cudaEventDestroy(e);
    cudaStreamDestroy(stream); cudaStreamDestroy(comm_stream);
  }

  inline size_t fidx(int ix, int iy, int iz) const { return (size_t)(ix + XO) + (size_t)PX * ((size_t)iy + (size_t)PY * iz); }
  inline size_t midx(int ix, int iy, int iz) const {
    return (size_t)(ix - 1) + (size_t)MPX * ((size_t)(iy - 1) + (size_t)N[1] * (iz - 1));
  }

  void set_pml_sigma(int group, int axis, const void* s, int len) override {
    if (finalized) throw std::string("khr_set_pml_sigma after khr_finalize_plan");
    int Ng = g.n[axis];
    if (len != 2 * Ng + 1) throw std::string("sigma profile length must be 2N+1");
    const T* sp = (const T*)s;
    int nl = N[axis];
    int off = (axis == 2) ? g.z_start - 1 : 0;
    std::vector<T>& v = h_sig[group][axis];
    v.assign((size_t)nl, T(0));
    // kernels sample sigma[2i-1] (1-based) for cell i (Helpers.jl:277)
    for (int i = 1; i <= nl; ++i) v[i - 1] = sp[2 * (i + off) - 2];
    have_sigma[group][axis] = true;
  }
  // DataStructures.jl:737-739: Δx/Δy/Δz may be vectors (one spacing per cell); the kernels then use
  // inv(Δ[i]) of the updated cell (Helpers.jl:283-291).  `d`: N global values of the context dtype.
  void set_grid_spacing(int axis, const void* d, int len) override {
    if (finalized) throw std::string("khr_set_grid_spacing after khr_finalize_plan");
    if (len != g.n[axis]) throw std::string("grid spacing vector must have one entry per cell of the axis");
    const T* sp = (const T*)d;
    const int nl = N[axis], off = (axis == 2) ? g.z_start - 1 : 0;
    const int cap = round_up(nl, 4) + 8;
    std::vector<T> inv((size_t)cap, T(1) / dl[axis]);
    for (int i = 0; i < nl; ++i) {
      if (!(sp[i + off] > T(0))) throw std::string("grid spacing must be positive");
      inv[(size_t)i] = T(1) / sp[i + off];   // inv(Δ[i])
    }
    if (!idv[axis]) idv[axis] = dalloc((size_t)cap, false);
    CUDA_OK(cudaMemcpyAsync(idv[axis], inv.data(), (size_t)cap * sizeof(T), cudaMemcpyHostToDevice, stream));
    CUDA_OK(cudaStreamSynchronize(stream));
    nonuniform = true;
  }
  // ---- geometry on the device (geom_kernels.cuh; Geometry.jl:150-246, 450-605, 795-972) --------
  // coordinate of local/global cell i on a component grid, the same expression as the kernels
  static double gcoord(double origin, int i, T d) { return origin + (double)((T)(i - 1) * d); }
  void geometry_rasterize(const khr_object* objs, int nobj, int kinds_mask, int smoothing, const double* origins18,
                          int64_t* smoothed3) override {
    if (finalized) throw std::string("khr_geometry_rasterize after khr_finalize_plan");
    if (nobj < 1) throw std::string("khr_geometry_rasterize: no objects");
    if (smoothing < 0 || smoothing > 2) throw std::string("khr_geometry_rasterize: smoothing must be 0 (none), 1 (volume averaging) or 2 (anisotropic)");
    if (nonuniform) throw std::string("khr_geometry_rasterize: non-uniform grids are rasterised by the caller");
    std::vector<GeomObj> h((size_t)nobj);
    for (int q = 0; q < nobj; ++q) {
      const khr_object& in = objs[q];
      GeomObj& o = h[(size_t)q];
      memset(&o, 0, sizeof(o));
      o.kind = in.kind;
      for (int k = 0; k < 3; ++k) o.c[k] = in.center[k];
      if (in.kind == KHR_SHAPE_SPHERE) {
        if (!(in.size[0] > 0)) throw std::string("khr_geometry_rasterize: sphere radius must be positive");
        o.r[0] = in.size[0];
        for (int k = 0; k < 3; ++k) { o.bmin[k] = o.c[k] - o.r[0]; o.bmax[k] = o.c[k] + o.r[0]; }   // bounds(::Sphere)
      } else if (in.kind == KHR_SHAPE_CUBOID) {
        bool ident = true;
        for (int k = 0; k < 9; ++k) ident = ident && in.axes[k] == 0.0;
        for (int k = 0; k < 3; ++k) {
          if (!(in.size[k] > 0)) throw std::string("khr_geometry_rasterize: cuboid sizes must be positive");
          o.r[k] = in.size[k] / 2;                                  // Cuboid(c, d, axes): r = d / 2
          double nr = 0;
          for (int j = 0; j < 3; ++j) { o.ax[3 * k + j] = ident ? (j == k ? 1.0 : 0.0) : in.axes[3 * k + j]; nr += o.ax[3 * k + j] * o.ax[3 * k + j]; }
          nr = std::sqrt(nr);
          if (!(nr > 0)) throw std::string("khr_geometry_rasterize: zero cuboid axis");
          for (int j = 0; j < 3; ++j) o.ax[3 * k + j] /= nr;
        }
        for (int a = 0; a < 3; ++a)
          for (int b = a + 1; b < 3; ++b) {
            double dot = 0;
            for (int j = 0; j < 3; ++j) dot += o.ax[3 * a + j] * o.ax[3 * b + j];
            if (std::fabs(dot) > 1e-12) throw std::string("khr_geometry_rasterize: cuboid axes must be orthogonal");
          }
        for (int i = 0; i < 3; ++i) {   // bounds(::Cuboid): c -+ sum_j |axis_j[i]| r_j
          double m = 0;
          for (int j = 0; j < 3; ++j) m += std::fabs(o.ax[3 * j + i]) * o.r[j];
          o.bmin[i] = o.c[i] - m; o.bmax[i] = o.c[i] + m;
        }
      } else if (in.kind == KHR_SHAPE_CYLINDER) {
        // Cylinder(c, r, h, a): size = {radius, height}, axes[0..2] = axis (normalised here);
        // bounds: c -+ (h/2 |a_i| + r sqrt(1 - a_i^2))
        if (!(in.size[0] > 0) || !(in.size[1] > 0)) throw std::string("khr_geometry_rasterize: cylinder radius and height must be positive");
        o.r[0] = in.size[0]; o.r[1] = in.size[1] / 2;
        double nr = std::sqrt((in.axes[0] * in.axes[0] + in.axes[1] * in.axes[1]) + in.axes[2] * in.axes[2]);
        if (!(nr > 0)) throw std::string("khr_geometry_rasterize: zero cylinder axis");
        for (int j = 0; j < 3; ++j) o.ax[j] = in.axes[j] / nr;
        for (int i = 0; i < 3; ++i) {
          const double m = o.r[1] * std::fabs(o.ax[i]) + o.r[0] * std::sqrt(std::max(0.0, 1.0 - o.ax[i] * o.ax[i]));
          o.bmin[i] = o.c[i] - m; o.bmax[i] = o.c[i] + m;
        }
      } else {
        throw std::string("khr_geometry_rasterize: shape kind must be KHR_SHAPE_SPHERE, KHR_SHAPE_CUBOID or KHR_SHAPE_CYLINDER");
      }
      for (int k = 0; k < 3; ++k) {
        o.val[0][k] = (double)(T)in.eps_inv[k]; o.val[1][k] = (double)(T)in.mu_inv[k];
        o.val[2][k] = (double)(T)in.sigma_d[k]; o.val[3][k] = (double)(T)in.sigma_b[k];
      }
    }
    GeomObj* d_objs = nullptr;
    int* d_owner = nullptr;
    PaintItem* d_items = nullptr;
    unsigned long long* d_cnt3 = nullptr;
    const size_t nown = (size_t)MPX * N[1] * (size_t)(N[2] + 2);
    auto cleanup = [&]() { cudaFree(d_objs); cudaFree(d_owner); cudaFree(d_items); cudaFree(d_cnt3); };
    try {
      CUDA_OK(cudaMalloc((void**)&d_objs, sizeof(GeomObj) * (size_t)nobj));
      CUDA_OK(cudaMalloc((void**)&d_owner, sizeof(int) * nown));
      CUDA_OK(cudaMalloc((void**)&d_cnt3, sizeof(unsigned long long) * 3));
      CUDA_OK(cudaMemsetAsync(d_cnt3, 0, sizeof(unsigned long long) * 3, stream));
      CUDA_OK(cudaMemcpyAsync(d_objs, h.data(), sizeof(GeomObj) * (size_t)nobj, cudaMemcpyHostToDevice, stream));
      size_t items_cap = 0;
      for (int c = 0; c < 6; ++c) {
        const int grp = c < 3 ? 1 : 0, d = c % 3;             // material group index: [0] = H side, [1] = E side
        const int perm_kind = c < 3 ? KHR_MAT_EPS_INV : KHR_MAT_MU_INV, sig_kind = c < 3 ? KHR_MAT_SIGMA_D : KHR_MAT_SIGMA_B;
        const bool want_perm = (kinds_mask >> perm_kind) & 1, want_sig = (kinds_mask >> sig_kind) & 1;
        if (!want_perm && !want_sig) continue;
        GeomGrid gg;
        for (int a = 0; a < 3; ++a) { gg.origin[a] = origins18[3 * c + a]; gg.n[a] = N[a]; }
        gg.nzg = g.n[2]; gg.z_off = g.z_start - 1; gg.mpx = MPX;
        // bounding-box index ranges on this grid: first cell with coord >= bmin .. last with coord <= bmax
        // (searchsortedfirst / searchsortedlast, Geometry.jl:188-194), clamped to the cells this slab
        // stores (z: one ghost plane either side, inside the global grid)
        std::vector<PaintItem> items;
        const int CH = 8192;
        for (int q = 0; q < nobj; ++q) {
          const GeomObj& o = h[(size_t)q];
          int lo[3], hi[3];
          bool empty = false;
          for (int a = 0; a < 3; ++a) {
            const int off = a == 2 ? gg.z_off : 0;
            const int cmin = a == 2 ? std::max(1 - off, 0) : 1;
            const int cmax = a == 2 ? std::min(g.n[2] - off, N[2] + 1) : N[a];
            const double dd = (double)dl[a];
            // estimate in double (objects like a 1e9-wide background box overflow an int), then clamp
            const double e0 = std::floor((o.bmin[a] - gg.origin[a]) / dd) + 1 - off - 2;
            const double e1 = std::ceil((o.bmax[a] - gg.origin[a]) / dd) + 1 - off + 2;
            int i0 = (int)std::min(std::max(e0, (double)cmin), (double)cmax + 1);
            int i1 = (int)std::max(std::min(e1, (double)cmax), (double)cmin - 1);
            while (i0 <= i1 && !(gcoord(gg.origin[a], i0 + off, dl[a]) >= o.bmin[a])) ++i0;
            while (i1 >= i0 && !(gcoord(gg.origin[a], i1 + off, dl[a]) <= o.bmax[a])) --i1;
            lo[a] = i0; hi[a] = i1;
            if (i0 > i1) empty = true;
          }
          if (empty) continue;
          const long long nv = (long long)(hi[0] - lo[0] + 1) * (hi[1] - lo[1] + 1) * (hi[2] - lo[2] + 1);
          for (long long f = 0; f < nv; f += CH) {
            PaintItem it;
            it.obj = q;
            for (int a = 0; a < 3; ++a) { it.lo[a] = lo[a]; it.n[a] = hi[a] - lo[a] + 1; }
            it.first = f; it.count = (int)std::min<long long>(CH, nv - f);
            items.push_back(it);
          }
        }
        CUDA_OK(cudaMemsetAsync(d_owner, 0x7f, sizeof(int) * nown, stream));
        if (!items.empty()) {
          if (items.size() > items_cap) {
            CUDA_OK(cudaStreamSynchronize(stream));
            cudaFree(d_items); d_items = nullptr;
            items_cap = items.size() * 2;
            CUDA_OK(cudaMalloc((void**)&d_items, sizeof(PaintItem) * items_cap));
          }
          CUDA_OK(cudaMemcpyAsync(d_items, items.data(), sizeof(PaintItem) * items.size(), cudaMemcpyHostToDevice, stream));
          geom_paint_kernel<T><<<(unsigned)items.size(), 256, 0, stream>>>(d_objs, d_items, gg, dl[0], dl[1], dl[2], d_owner);
          CUDA_OK(cudaGetLastError());
          CUDA_OK(cudaStreamSynchronize(stream));   // `items` is reused by the next grid
        }
        const long long nvox = (long long)N[0] * N[1] * N[2];
        const unsigned fb = (unsigned)std::min<long long>((nvox + 255) / 256, 148 * 32);
        if (want_perm) {
          if (!m_arr[grp][d]) m_arr[grp][d] = dalloc(msize);
          geom_fill_kernel<T><<<fb, 256, 0, stream>>>(d_objs, d_owner, gg, perm_kind, d, m_arr[grp][d]);
        }
        if (want_sig) {
          if (!sigM[grp][d]) sigM[grp][d] = dalloc(msize);
          geom_fill_kernel<T><<<fb, 256, 0, stream>>>(d_objs, d_owner, gg, sig_kind, d, sigM[grp][d]);
          has_sd[grp] = true;
        }
        if (want_perm && c < 3 && smoothing != 0)
          geom_smooth_kernel<T><<<(unsigned)((nvox + 255) / 256), 256, 0, stream>>>(d_objs, nobj, d_owner, gg, dl[0], dl[1], dl[2], d, smoothing,
                                                                                  m_arr[grp][d], d_cnt3 + d);
        CUDA_OK(cudaGetLastError());
        launches += 3;
      }
      unsigned long long cnt[3] = {0, 0, 0};
      CUDA_OK(cudaMemcpyAsync(cnt, d_cnt3, sizeof(cnt), cudaMemcpyDeviceToHost, stream));
      CUDA_OK(cudaStreamSynchronize(stream));
      if (smoothed3) for (int k = 0; k < 3; ++k) smoothed3[k] = (int64_t)cnt[k];
    } catch (...) {
      cudaStreamSynchronize(stream);
      cleanup();
      throw;
    }
    cleanup();
    // planner boxes of the conductive objects (cells, conservative: their bounding boxes on the cell-centre scale)
    for (int grp = 0; grp < 2; ++grp) {
      const int kind = grp == 1 ? KHR_MAT_SIGMA_D : KHR_MAT_SIGMA_B;
      if (!((kinds_mask >> kind) & 1)) continue;
      for (int q = 0; q < nobj; ++q) {
        const GeomObj& o = h[(size_t)q];
        if (o.val[kind][0] == 0 && o.val[kind][1] == 0 && o.val[kind][2] == 0) continue;
        int b[6];
        bool empty = false;
        for (int a = 0; a < 3; ++a) {
          const int off = a == 2 ? g.z_start - 1 : 0;
          // all six component grids lie within half a cell of the centre grid: pad by one cell
          const double org = origins18[3 * (grp == 1 ? 0 : 3) + a];
          const double e0 = std::floor((o.bmin[a] - org) / (double)dl[a]) + 1 - off - 1;
          const double e1 = std::ceil((o.bmax[a] - org) / (double)dl[a]) + 1 - off + 1;
          int i0 = (int)std::min(std::max(e0, 1.0), (double)N[a] + 1);
          int i1 = (int)std::max(std::min(e1, (double)N[a]), 0.0);
          b[a] = i0; b[3 + a] = i1;
          if (i0 > i1) empty = true;
        }
        if (!empty) box_union(sd_box[grp], b);
      }
    }
  }
  // dense (Nx,Ny,Nz_local) copy of a per-voxel material array (tests, host-side consumers)
  void material_read(int kind, int comp, void* out) override {
    if (comp < 0 || comp > 2) throw std::string("component must be 0..2");
    const T* src = nullptr;
    if (kind == KHR_MAT_EPS_INV) src = m_arr[1][comp];
    else if (kind == KHR_MAT_MU_INV) src = m_arr[0][comp];
    else if (kind == KHR_MAT_SIGMA_D) src = sigM[1][comp];
    else if (kind == KHR_MAT_SIGMA_B) src = sigM[0][comp];
    else if (kind == KHR_MAT_CHI3) src = chi3;
    else throw std::string("unknown material kind");
    if (!src) throw std::string("khr_material_read: this material array does not exist");
    std::vector<T> hh(msize);
    CUDA_OK(cudaStreamSynchronize(stream));
    CUDA_OK(cudaMemcpy(hh.data(), src, msize * sizeof(T), cudaMemcpyDeviceToHost));
    T* o = (T*)out;
    for (int z = 1; z <= N[2]; ++z)
      for (int y = 1; y <= N[1]; ++y)
        memcpy(o + (size_t)N[0] * ((size_t)(y - 1) + (size_t)N[1] * (z - 1)), hh.data() + midx(1, y, z), N[0] * sizeof(T));
  }
  void set_material_scalar(int kind, double v) override {
    if (kind == KHR_MAT_EPS_INV) m_scalar[1] = (T)v;
    else if (kind == KHR_MAT_MU_INV) m_scalar[0] = (T)v;
    else throw std::string("scalar material must be eps_inv or mu_inv");
  }
  // dense (Nx,Ny,Nzl) host -> material layout device
  T* upload_material(const void* dense, int* box6, std::vector<uint8_t>* mask = nullptr) {
    const T* src = (const T*)dense;
    if (mask) mask->resize((size_t)N[0] * N[1] * N[2], 0);
    std::vector<T> h(msize + 64, T(0));
    int b[6] = {1 << 30, 1 << 30, 1 << 30, 0, 0, 0};
    for (int z = 1; z <= N[2]; ++z)
      for (int y = 1; y <= N[1]; ++y) {
        const T* row = src + (size_t)N[0] * ((size_t)(y - 1) + (size_t)N[1] * (z - 1));
        T* dst = h.data() + midx(1, y, z);
        bool nz = false;
        int xl = 1 << 30, xh = 0;
        for (int x = 0; x < N[0]; ++x) {
          dst[x] = row[x];
          if (row[x] != T(0)) {
            nz = true; xl = std::min(xl, x + 1); xh = std::max(xh, x + 1);
            if (mask) (*mask)[(size_t)x + (size_t)N[0] * ((size_t)(y - 1) + (size_t)N[1] * (z - 1))] = 1;
          }
        }
        if (nz) {
          b[0] = std::min(b[0], xl); b[3] = std::max(b[3], xh);
          b[1] = std::min(b[1], y); b[4] = std::max(b[4], y);
          b[2] = std::min(b[2], z); b[5] = std::max(b[5], z);
        }
      }
    T* d = dalloc(msize, false);
    CUDA_OK(cudaMemcpyAsync(d, h.data(), (msize + 64) * sizeof(T), cudaMemcpyHostToDevice, stream));
    CUDA_OK(cudaStreamSynchronize(stream));
    if (box6) for (int q = 0; q < 6; ++q) box6[q] = b[q];
    return d;
  }
  static void box_union(int* a, const int* b) {
    if (b[0] > b[3]) return;
    if (a[0] > a[3]) { for (int q = 0; q < 6; ++q) a[q] = b[q]; return; }
    for (int q = 0; q < 3; ++q) { a[q] = std::min(a[q], b[q]); a[3 + q] = std::max(a[3 + q], b[3 + q]); }
  }
  void set_material_array(int kind, int comp, const void* dense) override {
    if (finalized) throw std::string("khr_set_material_array after khr_finalize_plan");
    if (comp < 0 || comp > 2) throw std::string("component must be 0..2");
    if (kind == KHR_MAT_CHI3) {
      if (comp != 0) throw std::string("chi3 is one array on the centre grid: component must be 0");
      chi3 = upload_material(dense, chi3_box, &chi3_mask);
      return;
    }
    int box[6];
    std::vector<uint8_t>* mk = nullptr;
    if (kind == KHR_MAT_SIGMA_D) mk = &sd_mask[1];
    if (kind == KHR_MAT_SIGMA_B) mk = &sd_mask[0];
    T* d = upload_material(dense, box, mk);
    if (kind == KHR_MAT_EPS_INV) m_arr[1][comp] = d;
    else if (kind == KHR_MAT_MU_INV) m_arr[0][comp] = d;
    else if (kind == KHR_MAT_SIGMA_D) { sigM[1][comp] = d; has_sd[1] = true; box_union(sd_box[1], box); }
    else if (kind == KHR_MAT_SIGMA_B) { sigM[0][comp] = d; has_sd[0] = true; box_union(sd_box[0], box); }
    else throw std::string("unknown material kind");
  }
  int pole_register(double omega0, double gamma, const void* sigma) override {
    if (finalized) throw std::string("khr_pole_register after khr_finalize_plan");
    if ((int)poles.size() >= 32) throw std::string("too many ADE poles (at most 32 per simulation)");
    // Susceptibility.jl:74-85 compute_ade_coefficients, Float64 then cast (Dispersive.jl:213-218)
    const double pi = 3.141592653589793;
    double dtd = (double)dt;
    double gpd = gamma * pi * dtd;
    double g1 = 1.0 - gpd, g1i = 1.0 / (1.0 + gpd);
    double w = (2 * pi) * omega0 * dtd;
    double w2 = w * w;
    double drude = gamma * (2 * pi) * dtd * dtd;
    Pole p;
    p.g1 = (T)g1; p.g1i = (T)g1i;
    if (omega0 == 0.0) { p.cp = T(2); p.cd = (T)drude; }
    else { p.cp = T(2) - (T)w2; p.cd = (T)w2; }
    pole_mask.emplace_back();
    p.sigma = upload_material(sigma, p.box, &pole_mask.back());
    for (int q = 0; q < 2; ++q)
      for (int d = 0; d < 3; ++d) p.P[q][d] = dalloc(msize);
    poles.push_back(p);
    return (int)poles.size() - 1;
  }
  int source_register(int comp, const int32_t* start, const int32_t* dims, const void* amp, int kind,
                      const double* tp) override {
    if (finalized) throw std::string("khr_source_register after khr_finalize_plan");
    Source s;
    s.comp = comp;
    size_t n = 1;
    for (int a = 0; a < 3; ++a) { s.s[a] = start[a]; s.d[a] = dims[a]; n *= (size_t)dims[a]; }
    s.s[2] -= g.z_start - 1;
    s.amp = dalloc(2 * n, false);
    CUDA_OK(cudaMemcpyAsync(s.amp, amp, 2 * n * sizeof(T), cudaMemcpyHostToDevice, stream));
    CUDA_OK(cudaStreamSynchronize(stream));
    s.ts.kind = kind;
    s.ts.fcen = (T)tp[0]; s.ts.width = (T)tp[1]; s.ts.peak = (T)tp[2]; s.ts.cutoff = (T)tp[3];
    sources.push_back(s);
    return (int)sources.size() - 1;
  }
  void source_set_amplitude(int id, double re, double im) override {
    if (id < 0 || id >= (int)sources.size()) throw std::string("bad source id");
    sources[id].ts.host_re = (T)re;
    sources[id].ts.host_im = (T)im;
  }
  int monitor_register(int comp, const int32_t* s, const int32_t* e, int nf, const double* f, int dec) override {
    if (finalized) throw std::string("khr_monitor_register after khr_finalize_plan");
    Monitor m;
    m.comp = comp;
    size_t n = 1;
    for (int a = 0; a < 3; ++a) { m.s[a] = s[a]; m.e[a] = e[a]; m.n[a] = e[a] - s[a] + 1; n *= (size_t)std::max(m.n[a], 0); }
    for (int k = 0; k < nf; ++k) m.freqs.push_back((T)f[k]);
    m.decimation = std::max(dec, 1);
    m.elems = n * (size_t)nf;
    m.M = dalloc(2 * m.elems);
    m.d_freqs = dalloc((size_t)nf, false);
    CUDA_OK(cudaMemcpyAsync(m.d_freqs, m.freqs.data(), nf * sizeof(T), cudaMemcpyHostToDevice, stream));
    CUDA_OK(cudaStreamSynchronize(stream));
    monitors.push_back(m);
    return (int)monitors.size() - 1;
  }

  // ---- planning -------------------------------------------------------------
  static bool boxes_hit(const int* b, int x0, int x1, int y0, int y1, int z0, int z1) {
    if (b[0] > b[3]) return false;
    return b[0] <= x1 && b[3] >= x0 && b[1] <= y1 && b[4] >= y0 && b[2] <= z1 && b[5] >= z0;
  }
  static int lx_index(int w) { return w >= 96 ? 2 : (w >= 48 ? 1 : 0); }
  // tile shape (0 / 1 / 2 = 32x32 / 64x16 / 128x8) that covers a w x h cell with the fewest tiles.  The wider shape
  // stays unless a narrower one saves at least 5 % of the tiles: on the DRAM-bound large grids full 128-voxel rows
  // stream slightly better (metalens 2048x2048x512: 82.4 vs 81.5 Gcells/s with 64x16 tiles that save 3 %)
  static int best_shape(int w, int h) {
    int best = 2; long long nb = -1;
    for (int si = 2; si >= 0; --si) {
      const int tw = 32 << si, th = 32 >> si;
      const long long n = (long long)((w + tw - 1) / tw) * ((h + th - 1) / th);
      if (nb < 0 || 20 * n <= 19 * nb) { nb = n; best = si; }
    }
    return best;
  }
  bool plan_fill = true;

  void finalize() override {
    if (finalized) throw std::string("khr_finalize_plan called twice");
    // the wrap kernels are plain launches between the chain kernels: keep stream semantics simple
    if (any_periodic() || in_pair || nonuniform) pdl = false;
    if (nonuniform) axis_spec = false;
    if (!pdl || g.nranks > 1) sweep = false;
    if (sweep) axis_spec = false;
    if (in_pair && chi3) throw std::string("chi3 with complex fields is not supported (|E|^2 couples the real and imaginary parts)");
    // per-axis PML cell sets from both groups' profiles
    std::vector<char> pml[3];
    for (int a = 0; a < 3; ++a) {
      pml[a].assign((size_t)N[a] + 2, 0);
      for (int gq = 0; gq < 2; ++gq)
        if (have_sigma[gq][a])
          for (int i = 1; i <= N[a]; ++i)
            if (h_sig[gq][a][i - 1] != T(0)) pml[a][i] = 1;
      // longest run of non-PML cells separates the lower from the upper slab
      int best_s = 1, best_len = 0, cur_s = 1, cur_len = 0;
      bool any = false;
      for (int i = 1; i <= N[a]; ++i) {
        if (!pml[a][i]) {
          if (cur_len == 0) cur_s = i;
          ++cur_len;
          if (cur_len > best_len) { best_len = cur_len; best_s = cur_s; }
        } else { cur_len = 0; any = true; }
      }
      if (!any) { lo_end[a] = 0; hi_start[a] = N[a] + 1; }
      else if (best_len == 0) { lo_end[a] = N[a]; hi_start[a] = N[a] + 1; }
      else { lo_end[a] = best_s - 1; hi_start[a] = best_s + best_len; }
    }
    // slab compaction
    {
      int n4 = round_up(N[0], 4);
      int low = round_up(lo_end[0], 4);
      int hib = hi_start[0] <= N[0] ? 1 + 4 * ((hi_start[0] - 1) / 4) : n4 + 1;
      if (low >= hib - 1) { low = n4; hib = n4 + 1; }
      slab[0].lo_w = low; slab[0].hi_base = hib;
      int cx = low + (n4 - hib + 1);
      cxp = std::max(round_up(cx, 32), 32);
      slab_elems[0] = (size_t)cxp * N[1] * N[2];
      if (cx == 0) slab_elems[0] = 0;
      slab[1].lo_w = lo_end[1]; slab[1].hi_base = hi_start[1];
      cy = lo_end[1] + (N[1] - hi_start[1] + 1);
      slab_elems[1] = (size_t)MPX * cy * N[2];
      slab[2].lo_w = lo_end[2]; slab[2].hi_base = hi_start[2];
      cz = lo_end[2] + (N[2] - hi_start[2] + 1);
      slab_elems[2] = (size_t)MPX * N[1] * cz;
    }
    for (int gq = 0; gq < 2; ++gq)
      for (int d = 0; d < 3; ++d) {
        int nx = (d + 1) % 3;
        if (slab_elems[d]) W[gq][d] = dalloc(slab_elems[d]);
        if (slab_elems[nx]) U[gq][d] = dalloc(slab_elems[nx]);
        bool anyp = slab_elems[0] || slab_elems[1] || slab_elems[2];
        if (has_sd[gq] && anyp) Cst[gq][d] = dalloc(msize);
        if (has_sd[gq] && !sigM[gq][d]) {  // absent component == zero conductivity
          sigM[gq][d] = dalloc(msize);
        }
      }
    if (chi3) {
      // A Kerr voxel rebuilds E from the stored D every step (KernelAbstractions path,
      // ReferenceKernels.jl:468-512); inside the PML the D-eliminated cascade has no exact
      // counterpart for a non-linear correction
      for (int z = 1; z <= N[2]; ++z)
        for (int y = 1; y <= N[1]; ++y)
          for (int x = 1; x <= N[0]; ++x)
            if (chi3_mask[(size_t)(x - 1) + (size_t)N[0] * ((size_t)(y - 1) + (size_t)N[1] * (z - 1))] &&
                (pml[0][x] || pml[1][y] || pml[2][z]))
              throw std::string("chi3 != 0 inside the PML is not supported (Kerr media must end before the PML)");
    }
    if (!poles.empty() || chi3)
      for (int d = 0; d < 3; ++d) Dst[d] = dalloc(msize);
    // coefficient vectors
    for (int gq = 0; gq < 2; ++gq)
      for (int a = 0; a < 3; ++a) {
        int len = round_up(N[a], 4) + 8;
        std::vector<T> s((size_t)len, T(0)), o((size_t)len, T(1)), ip((size_t)len, T(1));
        if (have_sigma[gq][a])
          for (int i = 0; i < N[a]; ++i) {
            T v = h_sig[gq][a][i];
            s[i] = v; o[i] = T(1) - v; ip[i] = T(1) / (T(1) + v);
          }
        for (int q = 0; q < 3; ++q) coef[gq][a][q] = dalloc((size_t)len, false);
        CUDA_OK(cudaMemcpyAsync(coef[gq][a][0], s.data(), len * sizeof(T), cudaMemcpyHostToDevice, stream));
        CUDA_OK(cudaMemcpyAsync(coef[gq][a][1], o.data(), len * sizeof(T), cudaMemcpyHostToDevice, stream));
        CUDA_OK(cudaMemcpyAsync(coef[gq][a][2], ip.data(), len * sizeof(T), cudaMemcpyHostToDevice, stream));
        CUDA_OK(cudaStreamSynchronize(stream));
      }
    if (nonuniform)
      for (int a = 0; a < 3; ++a)
        if (!idv[a]) {  // a uniform axis of a non-uniform grid: constant vector
          const int cap = round_up(N[a], 4) + 8;
          std::vector<T> inv((size_t)cap, T(1) / dl[a]);
          idv[a] = dalloc((size_t)cap, false);
          CUDA_OK(cudaMemcpyAsync(idv[a], inv.data(), (size_t)cap * sizeof(T), cudaMemcpyHostToDevice, stream));
          CUDA_OK(cudaStreamSynchronize(stream));
        }
    build_source_slots();
    {
      // Automatic choice, measured on B200 (profiles/r02_tma_ab.txt): the persistent TMA half-step kernel wins
      // where the PML tiles carry a large share of a large slab (sphere 512^3 +1.9 %, dipole 500^3 +1.0 %); small
      // grids profit more from the three overlapping LDG launches (waveguide -2.8 %), nearly PML-free ones
      // (metalens, 9 % PML voxels) from the LDG interior kernel that already runs at the copy peak.
      double npml = 1.0;
      for (int a = 0; a < 3; ++a) {
        int k = 0;
        for (int i = 1; i <= N[a]; ++i) k += pml[a][i] ? 1 : 0;
        npml *= 1.0 - (double)k / N[a];
      }
      const double cells = (double)N[0] * N[1] * N[2];
      tma_on = tma_policy == 1 || (tma_policy < 0 && cells >= 3.0e7 && (1.0 - npml) >= 0.25);
    }
    if (nonuniform || pdl || axis_spec || sizeof(T) != 4) tma_on = false;
    build_tables();
    if (tma_on) build_tensor_maps();
    // monitors
    if (!monitors.empty()) {
      std::vector<MonDesc<T>> h(monitors.size());
      for (size_t q = 0; q < monitors.size(); ++q) {
        Monitor& m = monitors[q];
        MonDesc<T>& d = h[q];
        d.M = m.M; d.F = F[m.comp]; d.Fi = im ? im->F[m.comp] : nullptr; d.freqs = m.d_freqs; d.nf = (int)m.freqs.size();
        d.decimation = m.decimation; d.group = m.comp >= 3 ? 0 : 1;
        // this rank accumulates the planes it owns; the top rank also owns the
        // staggered extra plane Nz+1 (never updated, zero, but part of the box)
        int zlo = g.z_start, zhi = g.z_start + N[2] - 1;
        if (g.rank == g.nranks - 1) zhi += 1;
        if (g.rank == 0) zlo = std::min(zlo, 0);
        int s2 = std::max(m.s[2], zlo), e2 = std::min(m.e[2], zhi);
        m.local = (e2 >= s2) && m.n[0] > 0 && m.n[1] > 0;
        for (int a = 0; a < 2; ++a) { d.s[a] = m.s[a]; d.n[a] = m.n[a]; d.moff[a] = 0; d.mn[a] = m.n[a]; }
        d.mn[2] = m.n[2];
        d.s[2] = s2 - (g.z_start - 1);
        d.n[2] = m.local ? (e2 - s2 + 1) : 0;
        d.moff[2] = s2 - m.s[2];
      }
      d_mons = (MonDesc<T>*)dalloc((sizeof(MonDesc<T>) * h.size() + sizeof(T) - 1) / sizeof(T), false);
      CUDA_OK(cudaMemcpyAsync(d_mons, h.data(), sizeof(MonDesc<T>) * h.size(), cudaMemcpyHostToDevice, stream));
      mon_local_nz.clear();
      for (auto& d : h) mon_local_nz.push_back(d.n[2]);
      CUDA_OK(cudaStreamSynchronize(stream));
    }
    CUDA_OK(cudaStreamSynchronize(stream));
    finalized = true;
  }

  // one flux-accumulator slot per (component, cell) touched by a source of the group; sources
  // that overlap on the same component share the slot
  void build_source_slots() {
    for (int gq = 0; gq < 2; ++gq) {
      std::vector<std::pair<unsigned long long, std::pair<size_t, size_t>>> keys;  // (key, (source, voxel))
      for (size_t q = 0; q < sources.size(); ++q) {
        Source& s = sources[q];
        if ((s.comp >= 3) != (gq == 0)) continue;
        size_t n = (size_t)s.d[0] * s.d[1] * s.d[2];
        for (size_t v = 0; v < n; ++v) {
          long long x = s.s[0] + (long long)(v % s.d[0]);
          long long y = s.s[1] + (long long)((v / s.d[0]) % s.d[1]);
          long long z = s.s[2] + (long long)(v / ((size_t)s.d[0] * s.d[1])) + 4;  // local z may be <= 0 on other ranks
          unsigned long long key = ((unsigned long long)(s.comp % 3) << 60) | ((unsigned long long)z << 40) |
                                   ((unsigned long long)y << 20) | (unsigned long long)x;
          keys.push_back({key, {q, v}});
        }
      }
      if (keys.empty()) continue;
      std::sort(keys.begin(), keys.end());
      std::vector<std::vector<int>> slots(sources.size());
      for (size_t q = 0; q < sources.size(); ++q)
        if ((sources[q].comp >= 3) == (gq == 0)) slots[q].assign((size_t)sources[q].d[0] * sources[q].d[1] * sources[q].d[2], 0);
      int cur = -1;
      unsigned long long last = ~0ull;
      for (auto& k : keys) {
        if (k.first != last) { ++cur; last = k.first; }
        slots[k.second.first][k.second.second] = cur;
      }
      nslots[gq] = (size_t)cur + 1;
      Tsrc[gq] = dalloc(nslots[gq]);
      for (size_t q = 0; q < sources.size(); ++q) {
        if (slots[q].empty()) continue;
        size_t bytes = slots[q].size() * sizeof(int);
        sources[q].slot = (int*)dalloc((bytes + sizeof(T) - 1) / sizeof(T), false);
        CUDA_OK(cudaMemcpyAsync(sources[q].slot, slots[q].data(), bytes, cudaMemcpyHostToDevice, stream));
        CUDA_OK(cudaStreamSynchronize(stream));
      }
    }
  }

  // split [1..n] at the PML edges; `gran` aligns the cut points outward (x only)
  struct Range { int s, e; bool pml; };
  std::vector<Range> axis_ranges(int a, int gran) const {
    std::vector<Range> r;
    int n = N[a];
    int lo = lo_end[a], hi = hi_start[a];
    if (lo <= 0 && hi > n) { r.push_back({1, n, false}); return r; }
    int lo_cut = std::min(n, round_up(lo, gran));                      // last cell of the lower PML range
    int hi_cut = hi <= n ? 1 + gran * ((hi - 1) / gran) : n + 1;      // first cell of the upper PML range
    if (lo_cut >= hi_cut - 1) { r.push_back({1, n, true}); return r; }
    if (lo_cut >= 1) r.push_back({1, lo_cut, true});
    r.push_back({lo_cut + 1, hi_cut - 1, false});
    if (hi_cut <= n) r.push_back({hi_cut, n, true});
    return r;
  }

  // Bytes one work item moves in one half-step, two models (w = sizeof(T)):
  //  * own (roofline.achieved): the compulsory traffic of THIS implementation — every array the tile
  //    touches once: 9w per voxel (3 curl operands read, 3 fields read + written), +3w for the per-voxel
  //    constitutive arrays unless the tile is constant (flags bit 1: the loads are skipped), +4w per PML
  //    axis of the voxel (W of that axis and U of the previous component, each read + written; B/D are
  //    eliminated), conductivity arrays 3w (+6w C stage inside the PML) on the tiles that carry them,
  //    ADE: 10w per pole voxel (sigma, P^n read, P^{n-1} read, P^{n+1} written) + D 6w on dispersive /
  //    Kerr voxels (+1w chi3).  Neighbour planes / rows shared between tiles come from L2 and are not
  //    counted, so the figure is a lower bound of the DRAM traffic (ncu: within 4 % of it).
  //  * ref (bytes_vs_reference_model): SURVEY.md §8(d), what the reference's layout would move for the
  //    same voxels — 9w (+3w per-voxel material), +10w / 14w / 18w on 1 / 2 / 3-PML-axis voxels,
  //    +3w conductivity, ADE 10w per pole + fPD 3w + D 6w + fPD read 3w.
  void item_bytes(int gq, const std::vector<int>* pmlc, const WorkItem& it, double* own, double* ref) const {
    const double w = sizeof(T);
    const int x0 = it.x0, xw = it.xw, y0 = it.y0, yh = it.yh, z0 = it.z0, zn = it.zn;
    int64_t p1[3], p0[3];
    int lo[3] = {x0, y0, z0}, n[3] = {xw, yh, zn};
    for (int a = 0; a < 3; ++a) {
      p1[a] = pmlc[a][lo[a] + n[a] - 1] - pmlc[a][lo[a] - 1];
      p0[a] = n[a] - p1[a];
    }
    const int64_t cells = (int64_t)xw * yh * zn;
    const int64_t c1 = p1[0] * p0[1] * p0[2] + p0[0] * p1[1] * p0[2] + p0[0] * p0[1] * p1[2];
    const int64_t c2 = p1[0] * p1[1] * p0[2] + p1[0] * p0[1] * p1[2] + p0[0] * p1[1] * p1[2];
    const int64_t c3 = p1[0] * p1[1] * p1[2];
    const bool marr = m_arr[gq][0] != nullptr;
    double b_ref = (double)cells * (marr ? 12.0 : 9.0) * w + (double)c1 * 10 * w + (double)c2 * 14 * w + (double)c3 * 18 * w;
    double b_own = (double)cells * ((marr && !(it.flags & 2)) ? 12.0 : 9.0) * w + (double)(c1 + 2 * c2 + 3 * c3) * 4 * w;
    auto count_mask = [&](const std::vector<uint8_t>& mk) {
      int64_t k = 0;
      if (mk.empty()) return k;
      for (int z = z0; z < z0 + zn; ++z)
        for (int y = y0; y < y0 + yh; ++y) {
          const uint8_t* r = mk.data() + (size_t)N[0] * ((size_t)(y - 1) + (size_t)N[1] * (z - 1));
          for (int x = x0; x < x0 + xw; ++x) k += r[x - 1];
        }
      return k;
    };
    if (has_sd[gq] && boxes_hit(sd_box[gq], x0, x0 + xw - 1, y0, y0 + yh - 1, z0, z0 + zn - 1)) {
      b_ref += (double)count_mask(sd_mask[gq]) * 3 * w;
      b_own += (double)cells * 3 * w + (Cst[gq][0] ? (double)(c1 + c2 + c3) * 6 * w : 0.0);
    }
    if (gq == 1)
      for (size_t q = 0; q < poles.size(); ++q)
        if (boxes_hit(poles[q].box, x0, x0 + xw - 1, y0, y0 + yh - 1, z0, z0 + zn - 1)) {
          const int64_t k = count_mask(pole_mask[q]);
          b_ref += (double)k * (10 + (q == 0 ? 12 : 0)) * w;
          b_own += (double)k * (10 + (q == 0 ? 6 : 0)) * w;
        }
    if (gq == 1 && chi3 && boxes_hit(chi3_box, x0, x0 + xw - 1, y0, y0 + yh - 1, z0, z0 + zn - 1)) {
      const int64_t k = count_mask(chi3_mask);
      b_ref += (double)k * 7 * w;  // chi3 read + D read/write
      b_own += (double)k * 7 * w;
    }
    *own = b_own; *ref = b_ref;
  }
  void account_tables(const std::vector<int>* pmlc) {
    auto one = [&](Table& t, int gq) {
      t.cells = 0; t.alg_bytes = 0; t.ref_bytes = 0; t.uniform_items = 0;
      for (auto& it : t.items) {
        double a, r;
        item_bytes(gq, pmlc, it, &a, &r);
        t.cells += (int64_t)it.xw * it.yh * it.zn;
        t.alg_bytes += a; t.ref_bytes += r;
        t.uniform_items += (it.flags & 2) ? 1 : 0;
      }
    };
    for_tables([&](Table& t, int gq, int, int) { one(t, gq); });
    one(tma_tab[0], 0); one(tma_tab[1], 1);
    if (getenv("KHR_PLAN_DUMP")) {
      // planner statistics: plane iterations (a CTA spends one iteration per plane of an item, whatever the item's
      // footprint) and how full the 1024-voxel planes are
      auto dump = [&](const char* what, Table& t, int gq, int m) {
        if (t.items.empty()) return;
        long long iters = 0, vox = 0, hist[4] = {0, 0, 0, 0};
        for (auto& it : t.items) {
          iters += it.zn; vox += (long long)it.xw * it.yh * it.zn;
          const double f = (double)it.xw * it.yh / 1024.0;
          hist[f > 0.95 ? 3 : f > 0.7 ? 2 : f > 0.45 ? 1 : 0] += it.zn;
        }
        fprintf(stderr, "[plan] %s group %d class %d: %zu items, %lld plane iterations, fill %.3f (iterations at <=45%% / <=70%% / <=95%% / full: %lld %lld %lld %lld)\n",
                what, gq, m, t.items.size(), iters, (double)vox / (1024.0 * iters), hist[0], hist[1], hist[2], hist[3]);
      };
      for_tables([&](Table& t, int gq, int ph, int m) { if (ph == 1) dump("ldg", t, gq, m); });
      dump("tma", tma_tab[0], 0, -1); dump("tma", tma_tab[1], 1, -1);
    }
  }

  template <class F>
  void for_tables(F f) {
    for (int gq = 0; gq < 2; ++gq)
      for (int ph = 0; ph < 2; ++ph)
        for (int m = 0; m < NTAB; ++m) f(tab[gq][ph][m], gq, ph, m);
  }
  void collect_profile() {
    auto one = [&](Table& t) {
      for (size_t q = 0; q + 1 < t.ev_used; q += 2) {
        float ms = 0;
        if (cudaEventElapsedTime(&ms, t.ev[q], t.ev[q + 1]) == cudaSuccess) { t.total_ms += ms; t.nlaunch += 1; }
        else cudaGetLastError();
      }
      t.ev_used = 0;
    };
    for_tables([&](Table& t, int, int, int) { one(t); });
    one(sweep_tab);
    one(tma_tab[0]); one(tma_tab[1]); one(fuse_tab); one(fix_tab[0]); one(fix_tab[1]);
  }
  void set_profiling(int on) override {
    sync_all();
    collect_profile();
    profiling = on != 0;
    serial_prof = on == 3;   // mode 3: one stream, so that every kernel's event pair times it alone
    collect_halo();
    if (on == 2 || on == 3) {
      for_tables([&](Table& t, int, int, int) { t.total_ms = 0; t.nlaunch = 0; });
      sweep_tab.total_ms = 0; sweep_tab.nlaunch = 0;
      for (int gq = 0; gq < 2; ++gq) { tma_tab[gq].total_ms = 0; tma_tab[gq].nlaunch = 0; }
      fuse_tab.total_ms = 0; fuse_tab.nlaunch = 0;
      for (int gq = 0; gq < 2; ++gq) { fix_tab[gq].total_ms = 0; fix_tab[gq].nlaunch = 0; }
      halo_wait_ms = 0; halo_exchanges = 0;
    }
  }
  int kernel_stat(int idx, khr_kernel_stat* out) override {
    sync_all();
    collect_profile();
    int k = 0;
    for_tables([&](Table& t, int gq, int ph, int m) {
      if (t.items.empty()) return;
      if (k == idx && out) {
        memset(out, 0, sizeof(*out));
        static const char* mn[MBASE] = {"interior", "pml-x", "pml-y", "", "pml-z", "", "", "pml", "full", "full-nopml"};
        snprintf(out->name, sizeof(out->name), "step_kernel<%s,%s,%s,%s>%s", sizeof(T) == 4 ? "f32" : "f64",
                 gq == 0 ? "H" : "E", mn[m % MBASE], m >= MBASE ? "muniform" : (m_arr[gq][0] ? "marr" : "mscalar"),
                 ph == 0 ? "[boundary]" : "");
        out->launches = t.nlaunch;
        out->total_ms = t.total_ms;
        out->cells_per_launch = t.cells;
        out->alg_bytes_per_launch = t.alg_bytes;
        out->ref_model_bytes_per_launch = t.ref_bytes;
        out->ctas = (int64_t)t.items.size();
        out->uniform_ctas = t.uniform_items;
      }
      ++k;
    });
    for (int gq = 0; gq < 2; ++gq) {
      Table& t = fix_tab[gq];
      if (t.items.empty()) continue;
      if (k == idx && out) {
        memset(out, 0, sizeof(*out));
        snprintf(out->name, sizeof(out->name), "fixup_kernel<%s,%s,sources+ade>", sizeof(T) == 4 ? "f32" : "f64", gq == 0 ? "H" : "E");
        out->launches = t.nlaunch; out->total_ms = t.total_ms; out->cells_per_launch = t.cells;
        out->alg_bytes_per_launch = 0; out->ref_model_bytes_per_launch = 0;
        out->ctas = (int64_t)t.items.size(); out->uniform_ctas = t.uniform_items;
      }
      ++k;
    }
    if (tma_fuse && !fuse_tab.items.empty()) {
      if (k == idx && out) {
        memset(out, 0, sizeof(*out));
        snprintf(out->name, sizeof(out->name), "step_tma_kernel<f32,H+E,interior+pml,%s/%s>", m_arr[0][0] ? "marr" : "mscalar", m_arr[1][0] ? "marr" : "mscalar");
        out->launches = fuse_tab.nlaunch; out->total_ms = fuse_tab.total_ms; out->cells_per_launch = fuse_tab.cells;
        out->alg_bytes_per_launch = fuse_tab.alg_bytes; out->ref_model_bytes_per_launch = fuse_tab.ref_bytes;
        out->ctas = (int64_t)fuse_tab.items.size(); out->uniform_ctas = fuse_tab.uniform_items;
      }
      ++k;
    }
    for (int gq = 0; gq < 2 && !tma_fuse; ++gq) {
      Table& t = tma_tab[gq];
      if (t.items.empty()) continue;
      if (k == idx && out) {
        memset(out, 0, sizeof(*out));
        snprintf(out->name, sizeof(out->name), "halfstep_tma_kernel<f32,%s,interior+pml,%s>", gq == 0 ? "H" : "E", m_arr[gq][0] ? "marr" : "mscalar");
        out->launches = t.nlaunch; out->total_ms = t.total_ms; out->cells_per_launch = t.cells;
        out->alg_bytes_per_launch = t.alg_bytes; out->ref_model_bytes_per_launch = t.ref_bytes;
        out->ctas = (int64_t)t.items.size(); out->uniform_ctas = t.uniform_items;
      }
      ++k;
    }
    if (sweep && !sweep_tab.items.empty()) {
      if (k == idx && out) {
        memset(out, 0, sizeof(*out));
        snprintf(out->name, sizeof(out->name), "sweep_kernel<%s,H+E,interior+pml,%s/%s>", sizeof(T) == 4 ? "f32" : "f64",
                 m_arr[0][0] ? "marr" : "mscalar", m_arr[1][0] ? "marr" : "mscalar");
        out->launches = sweep_tab.nlaunch;
        out->total_ms = sweep_tab.total_ms;
        out->cells_per_launch = sweep_tab.cells;
        out->alg_bytes_per_launch = sweep_tab.alg_bytes;
        out->ref_model_bytes_per_launch = sweep_tab.ref_bytes;
        out->ctas = (int64_t)sweep_tab.items.size();
        out->uniform_ctas = sweep_tab.uniform_items;
      }
      ++k;
    }
    return k;
  }

  // split ranges at extra cut points (cell index where a new range starts)
  static std::vector<Range> split_ranges(const std::vector<Range>& in, std::vector<int> cuts) {
    std::sort(cuts.begin(), cuts.end());
    std::vector<Range> out;
    for (auto r : in) {
      int s = r.s;
      for (int c : cuts)
        if (c > s && c <= r.e) { out.push_back({s, c - 1, r.pml}); s = c; }
      out.push_back({s, r.e, r.pml});
    }
    return out;
  }

  void build_tables() {
    std::vector<Range> xr0 = axis_ranges(0, 32), yr0 = axis_ranges(1, 1), zr0 = axis_ranges(2, 1);
    // prefix counts of PML cells per axis (for the bytes model)
    std::vector<int> pmlc[3];
    for (int a = 0; a < 3; ++a) {
      pmlc[a].assign((size_t)N[a] + 1, 0);
      for (int i = 1; i <= N[a]; ++i) {
        bool on = false;
        for (int gg = 0; gg < 2; ++gg)
          if (have_sigma[gg][a] && h_sig[gg][a][i - 1] != T(0)) on = true;
        pmlc[a][i] = pmlc[a][i - 1] + (on ? 1 : 0);
      }
    }
    // boxes that need the full kernel (sources, sigma_D/B, poles); the ranges are cut at
    // their faces so that only the voxels inside them pay for the extras
    struct Box { int b[6]; bool src; };
    std::vector<Box> gboxes[2];
    std::vector<int> czv;  // z cuts are shared by both field groups: the z chunks they define are
                           // the unit of the H <-> E dependency counters (chain mode)
    for (int gq = 0; gq < 2; ++gq) {
      std::vector<Box>& boxes = gboxes[gq];
      for (auto& s : sources)
        if ((s.comp >= 3) == (gq == 0))
          boxes.push_back({{s.s[0], s.s[1], s.s[2], s.s[0] + s.d[0] - 1, s.s[1] + s.d[1] - 1, s.s[2] + s.d[2] - 1}, true});
      if (has_sd[gq] && sd_box[gq][0] <= sd_box[gq][3]) {
        Box b; for (int q = 0; q < 6; ++q) b.b[q] = sd_box[gq][q]; b.src = false; boxes.push_back(b);
      }
      if (gq == 1)
        for (auto& pl : poles)
          if (pl.box[0] <= pl.box[3]) { Box b; for (int q = 0; q < 6; ++q) b.b[q] = pl.box[q]; b.src = false; boxes.push_back(b); }
      if (gq == 1 && chi3 && chi3_box[0] <= chi3_box[3]) {
        Box b; for (int q = 0; q < 6; ++q) b.b[q] = chi3_box[q]; b.src = false; boxes.push_back(b);
      }
      for (auto& bx : boxes) { czv.push_back(bx.b[2]); czv.push_back(bx.b[5] + 1); }
    }
    if (g.nranks > 1) {
      // the plane that feeds the halo exchange gets its own thin range (boundary-first launch)
      if (rank_up() >= 0) czv.push_back(N[2]);
      if (rank_dn() >= 0) czv.push_back(2);
    }
    const std::vector<Range> zr = split_ranges(zr0, czv);
    int zseg;
    {
      // z segment length: enough CTAs to fill 148 SMs several times over, but long enough
      // to amortise the carried plane
      long long tiles_xy = 0;
      for (auto& X : xr0) {
        int lx = 8 << lx_index(X.e - X.s + 1);
        for (auto& Y : yr0) tiles_xy += (long long)((X.e - X.s) / (4 * lx) + 1) * ((Y.e - Y.s) / (CTA / lx) + 1);
      }
      // measured on B200 (profiles/r01_zseg_sweep.txt): 7-8 planes per CTA is the sweet spot
      zseg = (int)std::min<long long>(8, std::max<long long>(4, (tiles_xy * N[2] + 2367) / 2368));
      // the persistent TMA kernel has no per-tile ramp to amortise and balances better with short tiles whose
      // neighbours are in flight at the same time (halo rows / carried planes hit L2): 3-4 planes measured best
      // (profiles/r02_tma_ab.txt: sphere 512^3 62.4 / 64.2 / 64.8 / 64.4 / 63.6 Gcells/s at 8 / 2 / 3 / 4 / 6)
      if (tma_on) zseg = 4;
      if (const char* e = getenv("KHR_ZSEG")) zseg = std::max(1, atoi(e));
    }
    // Source / pole / Kerr voxels outside the PML: the tile runs as a plain interior tile and fixup_kernel rewrites those
    // voxels afterwards (step_kernels.cuh).  Not with conductive media (the sigma_D stage sits inside the cascade), not on
    // several ranks (the boundary plane would need its own fix-up before the halo send), not in the chain / sweep / fused /
    // graph modes.
    for (int gq = 0; gq < 2; ++gq)
      fix_ok[gq] = fixup_on && !has_sd[gq] && g.nranks == 1 && !pdl && !sweep && !tma_fuse && !graph_on;
    auto set_zmask = [&](WorkItem& it) {
      it.zmask = 0;
      for (int q = 0; q < it.zn && q < 31; ++q)
        if (pmlc[2][it.z0 + q] - pmlc[2][it.z0 + q - 1]) it.zmask |= 1 << q;
    };
    nchunk = 0;
    for (auto& Z : zr) nchunk += (Z.e - Z.s) / zseg + 1;
    std::vector<unsigned long long> chunk_cnt[2];
    for (int gq = 0; gq < 2; ++gq) chunk_cnt[gq].assign((size_t)nchunk, 0ull);
    for (int gq = 0; gq < 2; ++gq) {
      std::vector<Box>& boxes = gboxes[gq];
      int chunk = -1;
      int zseg_full = 3;   // measured (profiles/r02_mode2_ab.txt): 3 planes x half-height CTAs: uled 45.5 -> 46.3, waveguide 49.5 -> 50.5 Gcells/s
      if (const char* e = getenv("KHR_ZSEG_FULL")) zseg_full = std::max(1, atoi(e));
      full_split = 2;
      if (const char* e = getenv("KHR_FULL_SPLIT")) full_split = atoi(e) == 1 ? 1 : 2;
      if (const char* e = getenv("KHR_FULL_NOPML")) full_nopml = atoi(e) != 0;
      bool local_cuts = true;   // x / y cuts only in the z ranges a box reaches (a point source no longer slices every plane)
      if (const char* e = getenv("KHR_LOCAL_CUTS")) local_cuts = atoi(e) != 0;
      if (const char* e = getenv("KHR_PLAN_FILL")) plan_fill = atoi(e) != 0;
      for (auto& Z : zr) {
        // the z ranges are already cut at the z faces of every box, so a box either spans Z or misses it
        std::vector<int> cx, cyv;
        for (auto& bx : boxes) {
          if (local_cuts && (bx.b[2] > Z.e || bx.b[5] < Z.s)) continue;
          cx.push_back(1 + 32 * ((std::max(bx.b[0], 1) - 1) / 32));
          cx.push_back(1 + 32 * ((std::max(bx.b[3], 0) + 31) / 32));
          cyv.push_back(bx.b[1]); cyv.push_back(bx.b[4] + 1);
        }
        const std::vector<Range> xr = split_ranges(xr0, cx), yr = split_ranges(yr0, cyv);
        // (x range, y range) cells of this z range.  plan_fill (default): the y cuts of a box apply only to the x ranges
        // the box reaches, and every cell takes the tile shape that covers it with the fewest tiles — a CTA spends the
        // same time on a plane of a tile whatever part of its 1024 voxels is inside the cell (KHR_PLAN_DUMP=1 prints
        // the fill; profiles/r02_plan_fill_ab.txt)
        struct Cell { Range X, Y; int lxi; };
        std::vector<Cell> cells;
        if (!plan_fill) {
          for (auto& Y : yr)
            for (auto& X : xr) cells.push_back({X, Y, lx_index(X.e - X.s + 1)});
        } else {
          for (auto& X : xr) {
            std::vector<int> cy;
            for (auto& bx : boxes) {
              if (local_cuts && (bx.b[2] > Z.e || bx.b[5] < Z.s)) continue;
              if (bx.b[0] > X.e || bx.b[3] < X.s) continue;
              cy.push_back(bx.b[1]); cy.push_back(bx.b[4] + 1);
            }
            for (auto& Y : split_ranges(yr0, cy)) cells.push_back({X, Y, best_shape(X.e - X.s + 1, Y.e - Y.s + 1)});
          }
          std::stable_sort(cells.begin(), cells.end(), [](const Cell& a, const Cell& b) { return a.Y.s != b.Y.s ? a.Y.s < b.Y.s : a.X.s < b.X.s; });
        }
        for (int z0 = Z.s; z0 <= Z.e; z0 += zseg) {
          int zn = std::min(zseg, Z.e - z0 + 1);
          ++chunk;
          for (auto& cl : cells) {
              const Range &X = cl.X, &Y = cl.Y;
              int lxi = cl.lxi;
              int lx = 8 << lxi;
              int tw = 4 * lx, th = CTA / lx;
              for (int y0 = Y.s; y0 <= Y.e; y0 += th) {
                int yh = std::min(th, Y.e - y0 + 1);
                for (int x0 = X.s; x0 <= X.e; x0 += tw) {
                  int xw = std::min(tw, X.e - x0 + 1);
                  // Tiles that touch a source / conductivity / pole box run the heavy MODE 2 kernel
                  // (1 CTA per SM): they are cut into short z pieces so that a thin box still
                  // spreads over all SMs and does not become the critical path of the half-step.
                  auto emit = [&](int zs, int zc, int ys = -1, int yc = 0) {
                    if (ys < 0) { ys = y0; yc = yh; }
                    WorkItem it{x0, xw, ys, yc, zs, zc, 0, 3 + lxi};
                    it.chunk = chunk;
                    chunk_cnt[gq][(size_t)chunk] += 1;
                    bool extras = false;
                    for (auto& bx : boxes)
                      if (boxes_hit(bx.b, x0, x0 + xw - 1, ys, ys + yc - 1, zs, zs + zc - 1)) {
                        extras = true;
                        if (bx.src) it.flags |= 1;
                      }
                    int axm = (X.pml ? 1 : 0) | (Y.pml ? 2 : 0) | (Z.pml ? 4 : 0);
                    it.flags |= axm << 4;   // axes whose PML the tile touches (pml_tma.cuh loads only their U / W slabs)
                    set_zmask(it);
                    if (!(axm == 1 || axm == 2 || axm == 4) || !axis_spec) axm = axm ? 7 : 0;
                    // optional: MODE 2 tiles outside the PML on the AXM = 0 variant of the full kernel (the cascade folds
                    // away: 144-184 registers instead of 220-250, no U / W traffic)
                    int mode = extras ? ((axm == 0 && full_nopml && !nonuniform && !pdl) ? 9 : 8) : axm;
                    if (extras && axm == 0 && fix_ok[gq]) { mode = 0; it.flags |= 8; }   // interior tile + fix-up of its special voxels
                    int phase = 1;
                    if (g.nranks > 1) {
                      if (gq == 0 && rank_up() >= 0 && zs + zc - 1 == N[2]) phase = 0;
                      if (gq == 1 && rank_dn() >= 0 && zs == 1) phase = 0;
                    }
                    Table& tt = tab[gq][phase][mode];
                    tt.items.push_back(it);
                    double ab, rb;
                    item_bytes(gq, pmlc, it, &ab, &rb);
                    tt.cost.push_back(rb);   // launch-order heuristic only; the tables are accounted at the end
                  };
                  bool any_box = false;
                  for (auto& bx : boxes)
                    if (boxes_hit(bx.b, x0, x0 + xw - 1, y0, y0 + yh - 1, z0, z0 + zn - 1)) any_box = true;
                  if (!any_box) emit(z0, zn);
                  else {
                    // MODE 2 CTAs have 128 threads (half the rows of a tile): two of them fit an SM at ~250
                    // registers, two independent load -> compute -> store chains per SM instead of one
                    const int hrows = std::max(1, th / full_split);
                    for (int zs = z0; zs < z0 + zn; zs += zseg_full)
                      for (int ys = y0; ys < y0 + yh; ys += hrows)
                        emit(zs, std::min(zseg_full, z0 + zn - zs), ys, std::min(hrows, y0 + yh - ys));
                  }
                }
              }
            }
        }
      }
    }
    // Launch order: the hardware hands CTAs to free SM slots in table order, so the heaviest
    // tiles (most PML axes) go first and the last ones are cut into short z pieces: the
    // tail of a half-step then ends within a few microseconds on every SM.
    int tail_zn = 0, sort_items = 0;  // both measured: no gain / a loss (sorting breaks L2 locality)
    if (const char* e = getenv("KHR_TAIL_ZN")) tail_zn = atoi(e);
    if (const char* e = getenv("KHR_SORT_ITEMS")) sort_items = atoi(e);
    for_tables([&](Table& t, int, int, int m) {
      if (t.items.empty() || m % MBASE >= 8) return;
      const size_t n = t.items.size();
      std::vector<size_t> idx(n);
      for (size_t q = 0; q < n; ++q) idx[q] = q;
      if (sort_items) std::stable_sort(idx.begin(), idx.end(), [&](size_t a, size_t b) { return t.cost[a] > t.cost[b]; });
      const size_t slots = (size_t)148 * (m % MBASE == 0 ? 3 : 2);
      const size_t ntail = tail_zn > 0 ? std::min(n / 3, slots) : 0;
      std::vector<WorkItem> out;
      out.reserve(n + 4 * ntail);
      for (size_t q = 0; q < n; ++q) {
        const WorkItem& it = t.items[idx[q]];
        if (q + ntail < n || it.zn <= tail_zn) { out.push_back(it); continue; }
        for (int zs = it.z0; zs < it.z0 + it.zn; zs += tail_zn) {
          WorkItem p = it;
          p.z0 = zs; p.zn = std::min(tail_zn, it.z0 + it.zn - zs);
          set_zmask(p);
          out.push_back(p);
        }
      }
      t.items.swap(out);
      t.cost.clear();
    });
    // Zigzag (KHR_ZIGZAG=1, off): the H launches walk the z chunks upwards, the E launches downwards, so
    // that a half-step starts on the planes the previous half-step touched last — still in the 126 MB
    // L2 — instead of on the ones it touched first.  Only the order of the tiles inside the E tables
    // changes (by chunk, descending; order inside a chunk kept).  Measured: no gain (waveguide 49.6 ->
    // 49.2, sphere 61.5 -> 61.3 Gcells/s, profiles/r01_s4_ab_zigzag.txt).
    {
      int zigzag = 0;
      if (const char* e = getenv("KHR_ZIGZAG")) zigzag = atoi(e);
      if (zigzag && !pdl)
        for_tables([&](Table& t, int gq, int, int) {
          if (gq != 1 || t.items.size() < 2) return;
          std::stable_sort(t.items.begin(), t.items.end(), [](const WorkItem& a, const WorkItem& b) { return a.chunk > b.chunk; });
        });
    }
    // items per (group, z chunk) after any splitting: the targets of the dependency counters
    for (int gq = 0; gq < 2; ++gq) std::fill(chunk_cnt[gq].begin(), chunk_cnt[gq].end(), 0ull);
    for_tables([&](Table& t, int gq, int, int) {
      for (auto& it : t.items) chunk_cnt[gq][(size_t)it.chunk] += 1;
    });
    for (int gq = 0; gq < 2; ++gq) {
      size_t words = ((size_t)nchunk * sizeof(unsigned long long) + sizeof(T) - 1) / sizeof(T);
      d_cnt[gq] = (unsigned long long*)dalloc(words, false);
      d_done[gq] = (unsigned long long*)dalloc(words, true);
      CUDA_OK(cudaMemcpyAsync(d_cnt[gq], chunk_cnt[gq].data(), (size_t)nchunk * sizeof(unsigned long long), cudaMemcpyHostToDevice, stream));
      epochs[gq] = 0;
    }
    CUDA_OK(cudaStreamSynchronize(stream));
    bool uniform_tiles = true;
    if (const char* e = getenv("KHR_UNIFORM_TILES")) uniform_tiles = atoi(e) != 0;
    for_tables([&](Table& t, int gq, int, int) {
      if (t.items.empty()) return;
      size_t bytes = t.items.size() * sizeof(WorkItem);
      t.d = (WorkItem*)dalloc((bytes + sizeof(T) - 1) / sizeof(T), false);
      CUDA_OK(cudaMemcpyAsync(t.d, t.items.data(), bytes, cudaMemcpyHostToDevice, stream));
      if (uniform_tiles && m_arr[gq][0] && m_arr[gq][1] && m_arr[gq][2]) {
        // tiles with constant material skip the per-voxel loads (bit-identical results)
        classify_items_kernel<T><<<(unsigned)t.items.size(), 256, 0, stream>>>(t.d, m_arr[gq][0], m_arr[gq][1], m_arr[gq][2], MPX,
                                                                             (long long)MPX * N[1]);
        CUDA_OK(cudaGetLastError());
        CUDA_OK(cudaMemcpyAsync(t.items.data(), t.d, bytes, cudaMemcpyDeviceToHost, stream));
      }
    });
    CUDA_OK(cudaStreamSynchronize(stream));
    for (int gq = 0; gq < 2; ++gq) {
      fix_tab[gq] = Table();
      for (auto& it : tab[gq][1][0].items)
        if (it.flags & 8) fix_tab[gq].items.push_back(it);
      if (fix_tab[gq].items.empty()) continue;
      const size_t bytes = fix_tab[gq].items.size() * sizeof(WorkItem);
      fix_tab[gq].d = (WorkItem*)dalloc((bytes + sizeof(T) - 1) / sizeof(T), false);
      CUDA_OK(cudaMemcpyAsync(fix_tab[gq].d, fix_tab[gq].items.data(), bytes, cudaMemcpyHostToDevice, stream));
      for (auto& it : fix_tab[gq].items) { fix_tab[gq].cells += (int64_t)it.xw * it.yh * it.zn; fix_tab[gq].uniform_items += (it.flags & 2) ? 1 : 0; }
    }
    // move the constant-material tiles of the interior / PML classes into their own tables
    if (split_uniform && uniform_tiles && !nonuniform && !pdl) {
      for (int gq = 0; gq < 2; ++gq)
        for (int ph = 0; ph < 2; ++ph)
          for (int m = 0; m < 8; ++m) {   // not the "full" classes
            // only the general PML class: its MARR = 1 kernel spills at the 128-register cap, the
            // MARR = 2 one does not (E-PML launch 0.145 -> 0.136 ms on the waveguide); the interior
            // kernel gains nothing from a second launch (measured, profiles/r01_s4_ab_split_uniform.txt)
            if (m != 7) continue;
            Table& t = tab[gq][ph][m];
            Table& u = tab[gq][ph][MBASE + m];
            if (t.items.empty()) continue;
            std::vector<WorkItem> keep;
            for (auto& it : t.items) ((it.flags & 2) ? u.items : keep).push_back(it);
            if (u.items.empty()) continue;
            t.items.swap(keep);
            for (Table* q : {&t, &u}) {
              if (q->items.empty()) { q->d = nullptr; continue; }
              size_t bytes = q->items.size() * sizeof(WorkItem);
              q->d = (WorkItem*)dalloc((bytes + sizeof(T) - 1) / sizeof(T), false);
              CUDA_OK(cudaMemcpyAsync(q->d, q->items.data(), bytes, cudaMemcpyHostToDevice, stream));
            }
          }
      CUDA_OK(cudaStreamSynchronize(stream));
    }
    if (tma_on && sizeof(T) == 4 && !nonuniform && !pdl && !axis_spec) {
      // The persistent TMA kernel takes the interior tiles and the PML tiles of phase 1 in ONE launch per
      // half-step (no tail between the classes, no wave quantisation).  PML tiles whose per-voxel material
      // is not constant stay behind (their three material boxes have no stage slot next to six aux boxes).
      int tail_zn2 = 2;
      if (const char* e = getenv("KHR_TMA_TAIL")) tail_zn2 = atoi(e);
      for (int gq = 0; gq < 2; ++gq) {
        const bool marr = m_arr[gq][0] != nullptr;
        Table& out = tma_tab[gq];
        out = Table();
        std::vector<std::pair<long long, WorkItem>> all;
        for (int m : {MBASE + 7, 7, 0}) {
          Table& t = tab[gq][1][m];
          std::vector<WorkItem> keep;
          for (size_t q = 0; q < t.items.size(); ++q) {
            const WorkItem& it = t.items[q];
            const bool pmlt = ((it.flags >> 4) & 7) != 0;
            if (it.zn > 31 || (pmlt && marr && !(it.flags & 2))) { keep.push_back(it); continue; }
            all.push_back({((long long)it.chunk * 2 + (pmlt ? 0 : 1)) * (1ll << 32) + (long long)all.size(), it});
          }
          t.items.swap(keep);
          if (t.items.empty()) t.d = nullptr;
          else {
            const size_t bytes = t.items.size() * sizeof(WorkItem);
            t.d = (WorkItem*)dalloc((bytes + sizeof(T) - 1) / sizeof(T), false);
            CUDA_OK(cudaMemcpyAsync(t.d, t.items.data(), bytes, cudaMemcpyHostToDevice, stream));
          }
        }
        std::sort(all.begin(), all.end(), [](const std::pair<long long, WorkItem>& a, const std::pair<long long, WorkItem>& b) { return a.first < b.first; });
        // the last items are cut into short z pieces: every SM runs out of work within a few planes of the others
        const size_t n = all.size(), ntail = tail_zn2 > 0 ? std::min(n / 4, (size_t)num_sms_guess * 2) : 0;
        for (size_t q = 0; q < n; ++q) {
          const WorkItem& it = all[q].second;
          if (q + ntail < n || it.zn <= tail_zn2) { out.items.push_back(it); continue; }
          for (int zs = it.z0; zs < it.z0 + it.zn; zs += tail_zn2) {
            WorkItem pc = it;
            pc.z0 = zs; pc.zn = std::min(tail_zn2, it.z0 + it.zn - zs);
            set_zmask(pc);
            out.items.push_back(pc);
          }
        }
        if (!out.items.empty()) {
          const size_t bytes = out.items.size() * sizeof(WorkItem);
          out.d = (WorkItem*)dalloc((bytes + sizeof(T) - 1) / sizeof(T), false);
          CUDA_OK(cudaMemcpyAsync(out.d, out.items.data(), bytes, cudaMemcpyHostToDevice, stream));
          out.d_ctr = (unsigned int*)dalloc(64 / sizeof(T), true);
        }
      }
      CUDA_OK(cudaStreamSynchronize(stream));
    }
    account_tables(pmlc);   // cells and both byte models per table, after classification / splitting
    if (tma_fuse && (!tma_on || g.nranks > 1 || any_periodic() || in_pair || tma_tab[0].items.empty() || tma_tab[1].items.empty())) tma_fuse = false;
    if (tma_fuse) {
      fuse_tab = Table();
      std::vector<std::pair<long long, WorkItem>> all;
      std::vector<unsigned int> cnt((size_t)nchunk, 0u);
      for (int gq = 0; gq < 2; ++gq) {
        const Table& t = tma_tab[gq];
        for (size_t q = 0; q < t.items.size(); ++q) {
          WorkItem it = t.items[q];
          it.flags |= gq << 8;
          if (gq == 0) cnt[(size_t)it.chunk] += 1;
          // E tiles of chunk c go `lag` chunks behind the H tiles of chunk c (lag >= 1: they wait for H(c-1), H(c) only)
          all.push_back({(((long long)it.chunk + (gq == 1 ? fuse_lag : 0)) * 2 + gq) * (1ll << 32) + (long long)q, it});
        }
        fuse_tab.cells += t.cells; fuse_tab.alg_bytes += t.alg_bytes; fuse_tab.ref_bytes += t.ref_bytes; fuse_tab.uniform_items += t.uniform_items;
      }
      std::sort(all.begin(), all.end(), [](const std::pair<long long, WorkItem>& a, const std::pair<long long, WorkItem>& b) { return a.first < b.first; });
      for (auto& kv : all) fuse_tab.items.push_back(kv.second);
      const size_t bytes = fuse_tab.items.size() * sizeof(WorkItem);
      fuse_tab.d = (WorkItem*)dalloc((bytes + sizeof(T) - 1) / sizeof(T), false);
      CUDA_OK(cudaMemcpyAsync(fuse_tab.d, fuse_tab.items.data(), bytes, cudaMemcpyHostToDevice, stream));
      fuse_tab.d_ctr = (unsigned int*)dalloc(64 / sizeof(T), true);
      d_fuse_done = (unsigned long long*)dalloc(((size_t)nchunk * 8 + sizeof(T) - 1) / sizeof(T) + 8, true);
      d_fuse_cnt = (unsigned int*)dalloc(((size_t)nchunk * 4 + sizeof(T) - 1) / sizeof(T) + 8, false);
      CUDA_OK(cudaMemcpyAsync(d_fuse_cnt, cnt.data(), (size_t)nchunk * 4, cudaMemcpyHostToDevice, stream));
      CUDA_OK(cudaStreamSynchronize(stream));
      fuse_epoch = 0;
    }
    if (sweep) {
      // chunk-major: H tiles of chunk c (PML first: they are the longer ones), then its E tiles
      sweep_tab = Table();
      int sweep_lag = 0;
      if (const char* e = getenv("KHR_SWEEP_LAG")) sweep_lag = std::max(0, atoi(e));
      std::vector<std::pair<long long, WorkItem>> all;
      for (int gq = 0; gq < 2; ++gq)
        for (int m : {7, 0}) {
          Table& t = tab[gq][1][m];
          for (size_t q = 0; q < t.items.size(); ++q) {
            WorkItem it = t.items[q];
            it.flags |= (gq << 8) | ((m == 7 ? 1 : 0) << 9);
            // E tiles of chunk c go `lag` chunks behind the H tiles (lag 0: right after H(c))
            const long long key = ((((long long)it.chunk + (gq == 1 ? sweep_lag : 0)) * 2 + gq) * 2 + (m == 7 ? 0 : 1)) * (1ll << 32) + (long long)q;
            all.push_back({key, it});
          }
          sweep_tab.cells += t.cells; sweep_tab.alg_bytes += t.alg_bytes; sweep_tab.ref_bytes += t.ref_bytes; sweep_tab.uniform_items += t.uniform_items;
        }
      std::sort(all.begin(), all.end(), [](const std::pair<long long, WorkItem>& a, const std::pair<long long, WorkItem>& b) { return a.first < b.first; });
      for (auto& kv : all) sweep_tab.items.push_back(kv.second);
      if (!sweep_tab.items.empty()) {
        size_t bytes = sweep_tab.items.size() * sizeof(WorkItem);
        sweep_tab.d = (WorkItem*)dalloc((bytes + sizeof(T) - 1) / sizeof(T), false);
        CUDA_OK(cudaMemcpyAsync(sweep_tab.d, sweep_tab.items.data(), bytes, cudaMemcpyHostToDevice, stream));
        CUDA_OK(cudaStreamSynchronize(stream));
      }
    }
  }

  // ---- TMA-staged PML kernel (pml_tma.cuh) ----------------------------------
  // Tensor maps live in one device array: [kind][array][tile shape], kinds A (halo box of a field array),
  // F (owner box of a field array), M (constitutive arrays), W, U (slab arrays); shape 0/1/2 = 32x32 / 64x16 / 128x8.
  bool tma_on = false;
  int tma_policy = -1;     // KHR_TMA: 1 on, 0 off, unset = automatic (large slabs with a sizeable PML share)
  int tma_stages_req = 0;
  CUtensorMap* d_maps = nullptr;
  enum { MAP_A = 0, MAP_F = 18, MAP_M = 36, MAP_W = 54, MAP_U = 72, MAP_COUNT = 90 };
  int num_sms = 148;
  static constexpr int num_sms_guess = 148;
  typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                    const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  void build_tensor_maps() {
    if (sizeof(T) != 4) { tma_on = false; return; }
    EncodeTiledFn enc = nullptr;
    cudaDriverEntryPointQueryResult qres;
    CUDA_OK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", (void**)&enc, cudaEnableDefault, &qres));
    if (!enc || qres != cudaDriverEntryPointSuccess) throw std::string("cuTensorMapEncodeTiled is not available in this driver");
    cudaDeviceProp prop;
    CUDA_OK(cudaGetDeviceProperties(&prop, device));
    num_sms = prop.multiProcessorCount;
    std::vector<CUtensorMap> h((size_t)MAP_COUNT);
    memset(h.data(), 0, sizeof(CUtensorMap) * h.size());
    auto make = [&](int slot, T* ptr, uint64_t d0, uint64_t d1, uint64_t d2, bool halo) {
      if (!ptr || d0 == 0 || d1 == 0 || d2 == 0) return;
      for (int si = 0; si < 3; ++si) {
        const uint32_t tw = 32u << si, th = 32u >> si;
        cuuint64_t dims[3] = {d0, d1, d2};
        cuuint64_t strides[2] = {d0 * sizeof(T), d0 * d1 * sizeof(T)};
        cuuint32_t box[3] = {halo ? tw + 4 : tw, halo ? th + 1 : th, 1};
        cuuint32_t es[3] = {1, 1, 1};
        CUresult r = enc(&h[(size_t)slot + si], CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, ptr, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) throw std::string("cuTensorMapEncodeTiled failed with code ") + std::to_string((int)r);
      }
    };
    for (int c = 0; c < 6; ++c) {
      make(MAP_A + 3 * c, F[c], (uint64_t)PX, (uint64_t)PY, (uint64_t)PZ, true);
      make(MAP_F + 3 * c, F[c], (uint64_t)PX, (uint64_t)PY, (uint64_t)PZ, false);
    }
    for (int gq = 0; gq < 2; ++gq)
      for (int d = 0; d < 3; ++d) {
        make(MAP_M + 3 * (3 * gq + d), m_arr[gq][d], (uint64_t)MPX, (uint64_t)N[1], (uint64_t)N[2], false);
        // W[d] lives on the slab of axis d, U[d] on the slab of axis next(d)
        const uint64_t sd[3][3] = {{(uint64_t)cxp, (uint64_t)N[1], (uint64_t)N[2]}, {(uint64_t)MPX, (uint64_t)cy, (uint64_t)N[2]},
                                   {(uint64_t)MPX, (uint64_t)N[1], (uint64_t)cz}};
        const int nx = (d + 1) % 3;
        make(MAP_W + 3 * (3 * gq + d), W[gq][d], sd[d][0], sd[d][1], sd[d][2], false);
        make(MAP_U + 3 * (3 * gq + d), U[gq][d], sd[nx][0], sd[nx][1], sd[nx][2], false);
      }
    d_maps = (CUtensorMap*)dalloc((sizeof(CUtensorMap) * h.size() + sizeof(T) - 1) / sizeof(T) + 64, false);
    // cudaMalloc returns 256-byte aligned memory: every 128-byte map is 64-byte aligned as TMA requires
    CUDA_OK(cudaMemcpyAsync(d_maps, h.data(), sizeof(CUtensorMap) * h.size(), cudaMemcpyHostToDevice, stream));
    CUDA_OK(cudaStreamSynchronize(stream));
  }
  template <int GROUP, bool RAGGED>
  void launch_tma_variant(const TmaParams<T>& tp, int stages, cudaStream_t st) {
    if constexpr (sizeof(T) == 4) {
      const int smem = tma_smem_bytes(0, stages);
      static bool attr_done = false;
      if (!attr_done) {
        CUDA_OK(cudaFuncSetAttribute(pml_tma_kernel<T, GROUP, RAGGED>, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448));
        attr_done = true;
      }
      const int grid = std::min(num_sms, tp.nitems);
      pml_tma_kernel<T, GROUP, RAGGED><<<grid, TMA_THREADS, smem, st>>>(tp);
    } else {
      throw std::string("internal: the TMA kernel is Float32 only");
    }
  }
  template <int GROUP>
  void fill_tma(const StepParams<T>& p, Table& t, TmaParams<T>& tp) {
    memset(&tp, 0, sizeof(tp));
    for (int d = 0; d < 3; ++d) {
      const int ia = GROUP == 0 ? d : 3 + d, jf = GROUP == 0 ? 3 + d : d;
      tp.mapA[d] = d_maps + MAP_A + 3 * ia;
      tp.mapF[d] = d_maps + MAP_F + 3 * jf;
      tp.mapM[d] = d_maps + MAP_M + 3 * (3 * GROUP + d);
      tp.mapW[d] = d_maps + MAP_W + 3 * (3 * GROUP + d);
      tp.mapU[d] = d_maps + MAP_U + 3 * (3 * GROUP + d);
      tp.F[d] = p.F[d]; tp.W[d] = p.W[d]; tp.U[d] = p.U[d];
      tp.n[d] = p.n[d]; tp.slab[d] = p.slab[d]; tp.idl[d] = p.idl[d];
      tp.sg[d] = p.sg[d]; tp.om[d] = p.om[d]; tp.ip[d] = p.ip[d];
    }
    tp.plane = p.plane; tp.mplane = p.mplane; tp.px = p.px; tp.mpx = p.mpx;
    tp.cxp = p.cxp; tp.cy = p.cy; tp.cz = p.cz;
    tp.dt = p.dt; tp.m_inv = p.m_inv;
    tp.marr = m_arr[GROUP][0] != nullptr ? 1 : 0;
    tp.items = t.d; tp.nitems = (int)t.items.size();
    tp.counter = t.d_ctr;
    int stages = (232448 - 1024 - 2 * TMA_ENTRY) / tma_stage_bytes(0);
    stages = std::min(stages, TMA_MAXSTAGES);
    if (tma_stages_req > 0) stages = std::min(stages, tma_stages_req);
    tp.nstages = stages;
  }
  template <int GROUP>
  void launch_tma(const StepParams<T>& p, Table& t, cudaStream_t st) {
    TmaParams<T> tp;
    fill_tma<GROUP>(p, t, tp);
    if ((N[0] % 4) != 0) launch_tma_variant<GROUP, true>(tp, tp.nstages, st);
    else launch_tma_variant<GROUP, false>(tp, tp.nstages, st);
    ++launches;
  }
  template <bool RAGGED>
  void launch_fused_variant(const TmaParams<T>& th_, const TmaParams<T>& te_, const TmaFuse& fu) {
    if constexpr (sizeof(T) == 4) {
      static bool attr_done = false;
      if (!attr_done) {
        CUDA_OK(cudaFuncSetAttribute(step_tma_kernel<T, RAGGED>, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448));
        attr_done = true;
      }
      const int grid = std::min(num_sms, th_.nitems);
      step_tma_kernel<T, RAGGED><<<grid, TMA_THREADS, tma_smem_bytes(0, th_.nstages), stream>>>(th_, te_, fu);
    } else {
      throw std::string("internal: the TMA kernel is Float32 only");
    }
  }
  // one time step with the fused kernel: [H tiles that need step_kernel] [step_tma_kernel] [E tiles that need step_kernel]
  void fused_step(double t, double th) {
    StepParams<T> ph, pe;
    fill_params(ph, 0, t);
    fill_params(pe, 1, th);
    const bool mh = m_arr[0][0] != nullptr, me = m_arr[1][0] != nullptr;
    skip_tma_launch = true;
    launch_group<0>(ph, 1, mh);
    TmaParams<T> th_, te_;
    fill_tma<0>(ph, fuse_tab, th_);
    fill_tma<1>(pe, fuse_tab, te_);
    TmaFuse fu;
    fuse_epoch += 1;
    fu.done_h = d_fuse_done; fu.cnt_h = d_fuse_cnt; fu.epoch = fuse_epoch; fu.nchunk = nchunk; fu.err_flag = d_err;
    timed(fuse_tab, true);
    if ((N[0] % 4) != 0) launch_fused_variant<true>(th_, te_, fu); else launch_fused_variant<false>(th_, te_, fu);
    ++launches;
    timed(fuse_tab, false);
    launch_group<1>(pe, 1, me);
    skip_tma_launch = false;
    CUDA_OK(cudaGetLastError());
    for (auto& pl : poles) pl.cur = 1 - pl.cur;
    epochs[0] += 1; epochs[1] += 1;
  }

  // ---- stepping -------------------------------------------------------------
  template <int GROUP, int MODE, int AXM>
  void launch_mode(const StepParams<T>& p, int marr, int n, cudaStream_t st) {
    if (pdl) {
      cudaLaunchConfig_t cfg = {};
      cfg.gridDim = dim3((unsigned)n); cfg.blockDim = dim3(MODE == 2 ? CTA / full_split : CTA); cfg.dynamicSmemBytes = 0; cfg.stream = st;
      cudaLaunchAttribute at[1];
      at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
      at[0].val.programmaticStreamSerializationAllowed = 1;
      cfg.attrs = at; cfg.numAttrs = 1;
      if (marr) CUDA_OK(cudaLaunchKernelEx(&cfg, step_kernel<T, GROUP, MODE, 1, AXM>, p));
      else CUDA_OK(cudaLaunchKernelEx(&cfg, step_kernel<T, GROUP, MODE, 0, AXM>, p));
    } else if (nonuniform) {
      // non-uniform grids use the general (AXM = 7) kernels
      if constexpr (AXM == 7) {
        if (marr) step_kernel<T, GROUP, MODE, 1, 7, true><<<n, MODE == 2 ? CTA / full_split : CTA, 0, st>>>(p);
        else step_kernel<T, GROUP, MODE, 0, 7, true><<<n, MODE == 2 ? CTA / full_split : CTA, 0, st>>>(p);
      } else {
        throw std::string("internal: axis-specialised kernels have no non-uniform variant");
      }
    } else {
      if (marr == 2) {
        if constexpr (AXM == 7 && MODE < 2) step_kernel<T, GROUP, MODE, 2, 7><<<n, CTA, 0, st>>>(p);
        else throw std::string("internal: no tile-uniform variant of this kernel");
      }
      else if (marr) step_kernel<T, GROUP, MODE, 1, AXM><<<n, MODE == 2 ? CTA / full_split : CTA, 0, st>>>(p);
      else step_kernel<T, GROUP, MODE, 0, AXM><<<n, MODE == 2 ? CTA / full_split : CTA, 0, st>>>(p);
    }
    if (capturing && MODE == 2) {
      // the node just added is the only dependency the next operation on this stream would get
      cudaStreamCaptureStatus cs;
      const cudaGraphNode_t* deps = nullptr;
      size_t ndeps = 0;
      CUDA_OK(cudaStreamGetCaptureInfo(st, &cs, nullptr, nullptr, &deps, &ndeps));
      if (cs != cudaStreamCaptureStatusActive || ndeps != 1) throw std::string("internal: cannot identify the captured MODE 2 kernel node");
      DynNode dn;
      dn.node = deps[0]; dn.gq = GROUP; dn.items = p.items;
      dyn_nodes.push_back(dn);
    }
    ++launches;
  }
  // The kernels of a half-step phase are independent: the interior one runs on the main
  // stream, every PML class and the full one on side streams, joined before the next phase.
  template <int GROUP>
  void launch_group(StepParams<T>& p, int phase, bool marr) {
    bool forked = false;
    bool used[NTAB] = {};
    // the heaviest table stays on the main stream, so the critical path of consecutive
    // half-steps is plain stream order with no cross-stream hop; it is launched last
    int mmain = 0;
    double best = -1;
    for (int m = 0; m < NTAB; ++m)
      if (!tab[GROUP][phase][m].items.empty() && tab[GROUP][phase][m].alg_bytes > best) { best = tab[GROUP][phase][m].alg_bytes; mmain = m; }
    if (!main_heaviest) mmain = 0;
    int order[NTAB], no = 0;
    // class c = m % MBASE; the tile-uniform table of a class (m = MBASE + c) goes right before its remainder
    auto push_class = [&](int c) { if (MBASE + c != mmain) order[no++] = MBASE + c; if (c != mmain) order[no++] = c; };
    if (launch_order == 1) {        // PML classes, interior, full last
      for (int c = 7; c >= 1; --c) push_class(c);
      push_class(0); push_class(9); push_class(8);
    } else if (launch_order == 2) { // PML classes, full, interior
      for (int c = 7; c >= 1; --c) push_class(c);
      push_class(9); push_class(8); push_class(0);
    } else {                        // full, PML classes, interior
      for (int c = MBASE - 1; c >= 0; --c) push_class(c);
    }
    order[no++] = mmain;
    for (int oi = 0; oi < no; ++oi) {
      const int m = order[oi];
      Table& t = tab[GROUP][phase][m];
      if (t.items.empty()) continue;
      cudaStream_t st = (m == mmain || !multi_stream || serial_prof) ? stream : side[side_of(m)];
      if (st != stream && !forked) {
        CUDA_OK(cudaEventRecord(ev_fork, stream));
        forked = true;
      }
      if (st != stream) CUDA_OK(cudaStreamWaitEvent(st, ev_fork, 0));
      p.items = t.d;
      int n = (int)t.items.size();
      if (profiling) {
        if (t.ev_used + 2 > t.ev.size()) {
          for (int q = 0; q < 2; ++q) { cudaEvent_t e; CUDA_OK(cudaEventCreate(&e)); t.ev.push_back(e); }
        }
        CUDA_OK(cudaEventRecord(t.ev[t.ev_used], st));
      }
      const int mk = m >= MBASE ? 2 : (marr ? 1 : 0);
      switch (m % MBASE) {
        case 0: launch_mode<GROUP, 0, 7>(p, mk, n, st); break;
        case 1: launch_mode<GROUP, 1, 1>(p, mk, n, st); break;
        case 2: launch_mode<GROUP, 1, 2>(p, mk, n, st); break;
        case 4: launch_mode<GROUP, 1, 4>(p, mk, n, st); break;
        case 8: launch_mode<GROUP, 2, 7>(p, mk, n, st); break;
        case 9: launch_mode<GROUP, 2, 0>(p, mk, n, st); break;
        default: launch_mode<GROUP, 1, 7>(p, mk, n, st); break;
      }
      if (profiling) {
        CUDA_OK(cudaEventRecord(t.ev[t.ev_used + 1], st));
        t.ev_used += 2;
      }
      if (st != stream) { CUDA_OK(cudaEventRecord(ev_join[side_of(m)], st)); used[m] = true; }
    }
    const bool tma_here = phase == 1 && tma_on && !skip_tma_launch && !tma_tab[GROUP].items.empty();
    if (phase == 1 && !fix_tab[GROUP].items.empty() && !tma_here) {
      // the fix-up pass depends only on the two interior tables (its tiles ran there as plain interior tiles): it
      // follows them on a stream of theirs, next to the PML launches, instead of after the join of the whole group
      const int m0 = 0, m1 = MBASE;
      const bool has0 = !tab[GROUP][phase][m0].items.empty(), has1 = !tab[GROUP][phase][m1].items.empty();
      auto st_of = [&](int m) { return (m == mmain || !multi_stream || serial_prof) ? stream : side[side_of(m)]; };
      cudaStream_t s0 = st_of(m0), s1 = st_of(m1);
      const int mf = (has0 && s0 != stream) ? m0 : (has1 && s1 != stream) ? m1 : (has0 ? m0 : m1);
      cudaStream_t fs = (has0 || has1) ? st_of(mf) : stream;
      const int mo = mf == m0 ? m1 : m0;
      if ((mo == m0 ? has0 : has1) && st_of(mo) != fs) {
        if (st_of(mo) == stream) {
          CUDA_OK(cudaEventRecord(ev_fix, stream));
          CUDA_OK(cudaStreamWaitEvent(fs, ev_fix, 0));
        } else {
          CUDA_OK(cudaStreamWaitEvent(fs, ev_join[side_of(mo)], 0));
        }
      }
      launch_fixup<GROUP>(p, fs);
      if (fs != stream) { CUDA_OK(cudaEventRecord(ev_join[side_of(mf)], fs)); used[mf] = true; }
    }
    if (tma_here) {
      // the persistent half-step kernel goes last, on the main stream: the small launches above have their
      // CTAs placed first, its 148 CTAs take the SMs as they become free and claim work dynamically
      Table& t = tma_tab[GROUP];
      timed(t, true);
      launch_tma<GROUP>(p, t, stream);
      timed(t, false);
      // with the TMA kernel the fix-up tiles ran inside it: the pass follows it in stream order
      if (!fix_tab[GROUP].items.empty()) launch_fixup<GROUP>(p, stream);
    }
    for (int m = 0; m < NTAB; ++m)
      if (used[m]) CUDA_OK(cudaStreamWaitEvent(stream, ev_join[side_of(m)], 0));
    CUDA_OK(cudaGetLastError());
  }
  void sync_all() {
    CUDA_OK(cudaStreamSynchronize(stream));
    for (int q = 0; q < NSIDE; ++q) CUDA_OK(cudaStreamSynchronize(side[q]));
  }

  double time_now() const { return (double)((T)timestep * dt); }  // Simulation.jl:22 round_time

  void fill_params(StepParams<T>& p, int gq, double t_src) {
    memset(&p, 0, sizeof(p));
    for (int d = 0; d < 3; ++d) {
      p.A[d] = gq == 0 ? F[d] : F[3 + d];
      p.F[d] = gq == 0 ? F[3 + d] : F[d];
      p.m_arr[d] = m_arr[gq][d];
      p.sg[d] = coef[gq][d][0]; p.om[d] = coef[gq][d][1]; p.ip[d] = coef[gq][d][2];
      p.W[d] = W[gq][d]; p.U[d] = U[gq][d];
      p.slab[d] = slab[d];
      p.sigD[d] = sigM[gq][d];
      p.C[d] = Cst[gq][d];
      p.n[d] = N[d];
      p.idl[d] = T(1) / dl[d];  // inv(Δ) (Helpers.jl:283)
      p.idv[d] = idv[d];
    }
    p.plane = (long long)PX * PY;
    p.px = PX;
    p.dt = dt;
    p.m_inv = m_scalar[gq];
    p.mpx = MPX;
    p.mplane = (long long)MPX * N[1];
    p.cxp = cxp; p.cy = cy; p.cz = cz;
    // sources of this group: up to MAXSRC descriptors ride in the kernel parameters, more go through a device
    // table refreshed on the stream before the launch (Sources.jl:330-340 loops over any number of sources)
    p.nsrc = 0;
    std::vector<SrcDesc<T>>& hs = h_src_ext[gq];
    hs.clear();
    for (auto& s : sources) {
      if ((s.comp >= 3) != (gq == 0)) continue;
      SrcDesc<T> d;
      d.amp = s.amp; d.slot = s.slot; d.comp = s.comp % 3;
      for (int a = 0; a < 3; ++a) { d.s[a] = s.s[a]; d.d[a] = s.d[a]; }
      T re = 0, im = 0;
      if (sources_active) eval_time_source(s.ts, t_src, &re, &im);
      d.an_re = re; d.an_im = im; d.ao_re = s.ao_re; d.ao_im = s.ao_im;
      s.ao_re = re; s.ao_im = im;
      hs.push_back(d);
    }
    p.nsrc = (int)hs.size();
    p.src_ext = nullptr;
    if (p.nsrc <= MAXSRC) {
      for (int q = 0; q < p.nsrc; ++q) p.src[q] = hs[(size_t)q];
    } else {
      if ((size_t)p.nsrc > src_ext_cap[gq]) {
        src_ext_cap[gq] = (size_t)p.nsrc;
        d_src_ext[gq] = (SrcDesc<T>*)dalloc((sizeof(SrcDesc<T>) * src_ext_cap[gq] + sizeof(T) - 1) / sizeof(T), false);
      }
      CUDA_OK(cudaMemcpyAsync(d_src_ext[gq], hs.data(), sizeof(SrcDesc<T>) * hs.size(), cudaMemcpyHostToDevice, stream));
      p.src_ext = d_src_ext[gq];
    }
    p.dep_on = pdl ? 1 : 0;
    p.nchunk = nchunk;
    p.done_mine = d_done[gq];
    p.done_other = d_done[1 - gq];
    p.cnt_other = d_cnt[1 - gq];
    p.epoch_other = epochs[1 - gq];   // every launch of the other group enqueued so far must have finished the chunk
    p.err_flag = d_err;
    p.npole = 0;
    for (int d = 0; d < 3; ++d) p.Dst[d] = (gq == 1) ? Dst[d] : nullptr;
    p.Tsrc = Tsrc[gq];
    p.chi3 = (gq == 1) ? chi3 : nullptr;
    p.pole_ext = nullptr;
    if (gq == 1) {
      std::vector<PoleDesc<T>>& hp = h_pole_ext;
      hp.clear();
      for (auto& pl : poles) {
        PoleDesc<T> d;
        d.sigma = pl.sigma;
        for (int c = 0; c < 3; ++c) { d.Pc[c] = pl.P[pl.cur][c]; d.Pp[c] = pl.P[1 - pl.cur][c]; }
        d.g1i = pl.g1i; d.g1 = pl.g1; d.cp = pl.cp; d.cd = pl.cd;
        hp.push_back(d);
      }
      p.npole = (int)hp.size();
      if (p.npole <= MAXPOLE) {
        for (int q = 0; q < p.npole; ++q) p.pole[q] = hp[(size_t)q];
      } else {
        // P^n / P^{n-1} swap roles every step, so the table is refreshed with the launch
        if (!d_pole_ext) d_pole_ext = (PoleDesc<T>*)dalloc((sizeof(PoleDesc<T>) * 32 + sizeof(T) - 1) / sizeof(T), false);
        CUDA_OK(cudaMemcpyAsync(d_pole_ext, hp.data(), sizeof(PoleDesc<T>) * hp.size(), cudaMemcpyHostToDevice, stream));
        p.pole_ext = d_pole_ext;
      }
    }
  }

  void update_sources_active(double t) {
    // Kernels.jl:27-35: sticky switch-off once t > last cutoff; CW never cuts off
    if (sources_mode == 0) { sources_active = false; return; }
    if (sources_mode == 1) { sources_active = true; return; }
    if (!sources_active || sources.empty()) return;
    double last = -1e300;
    for (auto& s : sources) {
      if (s.ts.kind == KHR_TIME_CW) return;
      last = std::max(last, (double)s.ts.cutoff);
    }
    if (t > last) sources_active = false;
  }

  void need_final() const {
    if (!finalized) throw std::string("khr_finalize_plan has not been called");
  }

  // send/recv of the two tangential components' boundary plane with the z neighbours
  void post_halo(int gq) {
    if (g.nranks <= 1) return;
    if (!comm) throw std::string("nranks > 1 but khr_comm_init was not called");
    size_t cnt = (size_t)PX * PY * sizeof(T);
    CUDA_OK(cudaEventRecord(ev_boundary, stream));
    CUDA_OK(cudaStreamWaitEvent(comm_stream, ev_boundary, 0));
    NCCL_OK(g_nccl.GroupStart());
    for (int c = 0; c < 2; ++c) {
      if (gq == 0) {
        // H: my top interior plane -> +z neighbour's lower ghost (E update reads H[k-1])
        T* f = F[3 + c];
        if (rank_up() >= 0) NCCL_OK(g_nccl.Send(f + (size_t)PX * PY * N[2], cnt, /*ncclChar*/ 0, rank_up(), comm, comm_stream));
        if (rank_dn() >= 0) NCCL_OK(g_nccl.Recv(f, cnt, 0, rank_dn(), comm, comm_stream));
      } else {
        // E: my bottom interior plane -> -z neighbour's upper ghost (H update reads E[k+1])
        T* f = F[c];
        if (rank_dn() >= 0) NCCL_OK(g_nccl.Send(f + (size_t)PX * PY, cnt, 0, rank_dn(), comm, comm_stream));
        if (rank_up() >= 0) NCCL_OK(g_nccl.Recv(f + (size_t)PX * PY * (N[2] + 1), cnt, 0, rank_up(), comm, comm_stream));
      }
    }
    NCCL_OK(g_nccl.GroupEnd());
    CUDA_OK(cudaEventRecord(ev_comm, comm_stream));
  }
  // profiling: CUDA events on the main stream right before and after the wait for the halo receive;
  // their distance is the time the next half-step could not start because the planes were not there yet
  std::vector<cudaEvent_t> halo_ev;
  size_t halo_ev_used = 0;
  double halo_wait_ms = 0;
  int64_t halo_exchanges = 0;
  void collect_halo() {
    for (size_t q = 0; q + 1 < halo_ev_used; q += 2) {
      float ms = 0;
      if (cudaEventElapsedTime(&ms, halo_ev[q], halo_ev[q + 1]) == cudaSuccess) { halo_wait_ms += ms; halo_exchanges += 1; }
      else cudaGetLastError();
    }
    halo_ev_used = 0;
  }
  void comm_stat(double* wait_ms, int64_t* exchanges) override {
    sync_all();
    collect_halo();
    if (wait_ms) *wait_ms = halo_wait_ms;
    if (exchanges) *exchanges = halo_exchanges;
  }
  void wait_halo() {
    if (g.nranks <= 1) return;
    if (profiling) {
      if (halo_ev_used + 2 > halo_ev.size())
        for (int q = 0; q < 2; ++q) { cudaEvent_t e; CUDA_OK(cudaEventCreate(&e)); halo_ev.push_back(e); }
      CUDA_OK(cudaEventRecord(halo_ev[halo_ev_used], stream));
    }
    CUDA_OK(cudaStreamWaitEvent(stream, ev_comm, 0));
    if (profiling) {
      CUDA_OK(cudaEventRecord(halo_ev[halo_ev_used + 1], stream));
      halo_ev_used += 2;
      if (halo_ev_used >= 4096) { CUDA_OK(cudaStreamSynchronize(stream)); collect_halo(); }
    }
  }

  // periodic axes: wrap the group's three components once its kernels are queued (the
  // reference calls exchange_halos! at the end of step_H_fused! / step_E_fused!, before the
  // DFT update).  x and y are always local to the slab; z is local only on a single rank,
  // otherwise the wrap travels with the halo exchange (ring, post_halo).
  void wrap_periodic(int gq) {
    T* f[3];
    for (int d = 0; d < 3; ++d) f[d] = gq == 0 ? F[3 + d] : F[d];
    const long long st[3] = {1, (long long)PX, (long long)PX * PY};
    for (int a = 0; a < 3; ++a) {
      if (!periodic[a]) continue;
      if (a == 2 && g.nranks > 1) continue;
      const int t1 = (a + 1) % 3, t2 = (a + 2) % 3;
      const long long cells = (long long)N[t1] * N[t2];
      dim3 grid((unsigned)((cells + 255) / 256), 3);
      wrap_kernel<T><<<grid, 256, 0, stream>>>(f[0], f[1], f[2], (long long)XO, st[a], st[t1], st[t2], N[a], N[t1], N[t2]);
      ++launches;
    }
    CUDA_OK(cudaGetLastError());
  }
  bool any_periodic() const { return periodic[0] || periodic[1] || periodic[2]; }
  // z neighbours of this rank (ring when z is periodic), -1 if none
  int rank_up() const { return g.rank < g.nranks - 1 ? g.rank + 1 : (periodic[2] && g.nranks > 1 ? 0 : -1); }
  int rank_dn() const { return g.rank > 0 ? g.rank - 1 : (periodic[2] && g.nranks > 1 ? g.nranks - 1 : -1); }

  void half_step(int gq, double t_src) {
    StepParams<T> p;
    fill_params(p, gq, t_src);
    half_step_launch(gq, p);
    if (gq == 1) for (auto& pl : poles) pl.cur = 1 - pl.cur;
    epochs[gq] += 1;
  }
  template <int GROUP>
  void launch_fixup(StepParams<T>& p, cudaStream_t st) {
    Table& t = fix_tab[GROUP];
    if (t.items.empty()) return;
    p.items = t.d;
    const int n = (int)t.items.size();
    const bool marr = m_arr[GROUP][0] != nullptr;
    int maxc = 1;
    for (auto& it : t.items) maxc = std::max(maxc, it.xw * it.yh * it.zn);
    const dim3 grid((unsigned)n, (unsigned)((maxc + 255) / 256));
    timed(t, true, st);
    if (nonuniform) {
      if (marr) fixup_kernel<T, GROUP, 1, true><<<grid, 256, 0, st>>>(p);
      else fixup_kernel<T, GROUP, 0, true><<<grid, 256, 0, st>>>(p);
    } else {
      if (marr) fixup_kernel<T, GROUP, 1, false><<<grid, 256, 0, st>>>(p);
      else fixup_kernel<T, GROUP, 0, false><<<grid, 256, 0, st>>>(p);
    }
    timed(t, false, st);
    ++launches;
  }
  void half_step_launch(int gq, StepParams<T>& p) {
    bool marr = m_arr[gq][0] != nullptr;
    if (marr && (!m_arr[gq][1] || !m_arr[gq][2])) throw std::string("per-voxel material needs all three components");
    if (g.nranks > 1) {
      // boundary planes first, ship them while the interior updates
      if (gq == 0) launch_group<0>(p, 0, marr); else launch_group<1>(p, 0, marr);
      post_halo(gq);
      if (gq == 0) launch_group<0>(p, 1, marr); else launch_group<1>(p, 1, marr);
      wait_halo();
    } else {
      if (gq == 0) launch_group<0>(p, 1, marr); else launch_group<1>(p, 1, marr);
    }
    if (any_periodic() && !in_pair) wrap_periodic(gq);
  }
  // ---- one step as a CUDA graph ------------------------------------------------------------------------------
  void graph_info(int64_t* kernels_per_graph, int64_t* replays) override {
    if (kernels_per_graph) *kernels_per_graph = step_gexec ? graph_kernels : 0;
    if (replays) *replays = graph_replays;
  }
  bool graph_usable() const {
    return graph_on && !pdl && !sweep && !tma_fuse && g.nranks == 1 && !in_pair && !profiling;
  }
  void graph_step(double t, double th) {
    StepParams<T> ph, pe;
    fill_params(ph, 0, t);
    fill_params(pe, 1, th);
    if (!step_gexec) {
      dyn_nodes.clear();
      const int64_t l0 = launches;
      capturing = true;
      cudaError_t e = cudaStreamBeginCapture(stream, cudaStreamCaptureModeThreadLocal);
      if (e != cudaSuccess) { capturing = false; throw std::string("cudaStreamBeginCapture: ") + cudaGetErrorString(e); }
      try {
        half_step_launch(0, ph);
        half_step_launch(1, pe);
      } catch (...) {
        capturing = false;
        cudaGraph_t junk = nullptr;
        cudaStreamEndCapture(stream, &junk);
        if (junk) cudaGraphDestroy(junk);
        throw;
      }
      capturing = false;
      CUDA_OK(cudaStreamEndCapture(stream, &step_graph));
      graph_kernels = launches - l0;
      launches = l0;
      CUDA_OK(cudaGraphInstantiate(&step_gexec, step_graph, 0));
      for (auto& dn : dyn_nodes) CUDA_OK(cudaGraphKernelNodeGetParams(dn.node, &dn.kp));
    }
    for (auto& dn : dyn_nodes) {
      StepParams<T> q = dn.gq == 0 ? ph : pe;
      q.items = dn.items;
      void* args[1] = {&q};
      cudaKernelNodeParams kp = dn.kp;
      kp.kernelParams = args;
      kp.extra = nullptr;
      CUDA_OK(cudaGraphExecKernelNodeSetParams(step_gexec, dn.node, &kp));
    }
    CUDA_OK(cudaGraphLaunch(step_gexec, stream));
    launches += graph_kernels;
    graph_replays += 1;
    for (auto& pl : poles) pl.cur = 1 - pl.cur;
    epochs[0] += 1; epochs[1] += 1;
  }
  // complex fields: the same half-step on the imaginary parts (own streams, no sources), then the
  // wrap-around copies of both parts with the Bloch phase on this context's stream
  void half_step_all(int gq, double t_src) {
    half_step(gq, t_src);
    if (!im) return;
    im->half_step(gq, t_src);
    launches += im->launches; im->launches = 0;
    CUDA_OK(cudaEventRecord(ev_pair_a, im->stream));
    CUDA_OK(cudaStreamWaitEvent(stream, ev_pair_a, 0));
    const long long st[3] = {1, (long long)PX, (long long)PX * PY};
    for (int a = 0; a < 3; ++a) {
      if (!periodic[a]) continue;
      if (a == 2 && g.nranks > 1) {
        // z ring over the ranks: the wrapped planes arrived with the halo exchange of both contexts (H: the top
        // plane of the last rank in the lower ghost of rank 0; E: the bottom plane of rank 0 in the upper ghost of
        // the last rank); what is left of Chunking.jl:1735-1764 is the phase on those two ghost planes
        const double kl = bloch_kl[2];
        const bool lower = gq == 0 && g.rank == 0, upper = gq == 1 && g.rank == g.nranks - 1;
        if ((lower || upper) && kl != 0.0) {
          const double pr = std::cos(lower ? -kl : kl), pi_ = std::sin(lower ? -kl : kl);
          if (!(pr == 1.0 && pi_ == 0.0)) {
            const long long off = lower ? 0 : (long long)PX * PY * (N[2] + 1);
            const long long cnt = (long long)PX * PY;
            for (int d = 0; d < 2; ++d) {    // the two tangential components that travel
              T* fr = (gq == 0 ? F[3 + d] : F[d]) + off;
              T* fi = (gq == 0 ? im->F[3 + d] : im->F[d]) + off;
              bloch_phase_kernel<T><<<(unsigned)((cnt + 255) / 256), 256, 0, stream>>>(fr, fi, cnt, pr, pi_);
              ++launches;
            }
          }
        }
        continue;
      }
      BlochWrapArgs<T> w;
      for (int d = 0; d < 3; ++d) { w.fr[d] = gq == 0 ? F[3 + d] : F[d]; w.fi[d] = gq == 0 ? im->F[3 + d] : im->F[d]; }
      const int t1 = (a + 1) % 3, t2 = (a + 2) % 3;
      w.base = XO; w.sa = st[a]; w.s1 = st[t1]; w.s2 = st[t2];
      w.n_axis = N[a]; w.n1 = N[t1]; w.n2 = N[t2];
      // phase_fwd = exp(i k L), phase_rev = exp(-i k L) (Chunking.jl:1746-1747)
      w.fwd_re = std::cos(bloch_kl[a]); w.fwd_im = std::sin(bloch_kl[a]);
      w.rev_re = std::cos(-bloch_kl[a]); w.rev_im = std::sin(-bloch_kl[a]);
      w.apply_fwd = !(w.fwd_re == 1.0 && w.fwd_im == 0.0);
      w.apply_rev = !(w.rev_re == 1.0 && w.rev_im == 0.0);
      const long long cells = (long long)N[t1] * N[t2];
      bloch_wrap_kernel<T><<<dim3((unsigned)((cells + 255) / 256), 3), 256, 0, stream>>>(w);
      ++launches;
    }
    CUDA_OK(cudaGetLastError());
    CUDA_OK(cudaEventRecord(ev_pair_b, stream));
    CUDA_OK(cudaStreamWaitEvent(im->stream, ev_pair_b, 0));
  }

  // one time step in sweep mode: [H tiles that need the full kernel] [sweep grid] [E tiles that need the
  // full kernel], all on one stream with programmatic dependent launch; ordering by the chunk counters
  template <int MH, int ME>
  void launch_sweep(const StepParams<T>& ph, const StepParams<T>& pe) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)sweep_tab.items.size()); cfg.blockDim = dim3(CTA); cfg.dynamicSmemBytes = 0; cfg.stream = stream;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    CUDA_OK(cudaLaunchKernelEx(&cfg, sweep_kernel<T, MH, ME>, ph, pe, (const WorkItem*)sweep_tab.d));
    ++launches;
  }
  void timed(Table& t, bool begin, cudaStream_t st = nullptr) {
    if (!profiling) return;
    if (!st) st = stream;
    if (begin) {
      if (t.ev_used + 2 > t.ev.size())
        for (int q = 0; q < 2; ++q) { cudaEvent_t e; CUDA_OK(cudaEventCreate(&e)); t.ev.push_back(e); }
      CUDA_OK(cudaEventRecord(t.ev[t.ev_used], st));
    } else {
      CUDA_OK(cudaEventRecord(t.ev[t.ev_used + 1], st));
      t.ev_used += 2;
    }
  }
  void sweep_step(double t, double th) {
    StepParams<T> ph, pe;
    fill_params(ph, 0, t);
    epochs[0] += 1;
    fill_params(pe, 1, th);
    const bool mh = m_arr[0][0] != nullptr, me = m_arr[1][0] != nullptr;
    if ((mh && (!m_arr[0][1] || !m_arr[0][2])) || (me && (!m_arr[1][1] || !m_arr[1][2])))
      throw std::string("per-voxel material needs all three components");
    Table& hf = tab[0][1][8];
    if (!hf.items.empty()) {
      ph.items = hf.d;
      timed(hf, true);
      launch_mode<0, 2, 7>(ph, mh ? 1 : 0, (int)hf.items.size(), stream);
      timed(hf, false);
    }
    if (!sweep_tab.items.empty()) {
      timed(sweep_tab, true);
      if (mh && me) launch_sweep<1, 1>(ph, pe);
      else if (mh) launch_sweep<1, 0>(ph, pe);
      else if (me) launch_sweep<0, 1>(ph, pe);
      else launch_sweep<0, 0>(ph, pe);
      timed(sweep_tab, false);
    }
    Table& ef = tab[1][1][8];
    if (!ef.items.empty()) {
      pe.items = ef.d;
      timed(ef, true);
      launch_mode<1, 2, 7>(pe, me ? 1 : 0, (int)ef.items.size(), stream);
      timed(ef, false);
    }
    CUDA_OK(cudaGetLastError());
    for (auto& pl : poles) pl.cur = 1 - pl.cur;
    epochs[1] += 1;
  }
  void step_h() override {
    need_final();
    double t = time_now();
    update_sources_active(t);
    half_step_all(0, t);
  }
  void step_e() override {
    need_final();
    double t = time_now() + (double)(dt / T(2));
    half_step_all(1, t);
  }
  void dft_update(int group, double time) override {
    need_final();
    if (monitors.empty()) return;
    const double two_pi = 2 * 3.141592653589793;
    const T tf = (T)(two_pi * time);  // Monitors.jl:323 complex_backend_number(im*2π*time), cast to T
    DftBatch<T> bt;
    bt.n = 0;
    int used = 0;
    long long maxcells = 0;
    auto flush = [&]() {
      if (bt.n == 0) return;
      dim3 grid((unsigned)((maxcells + 255) / 256), (unsigned)bt.n);
      dft_kernel<T><<<grid, 256, 0, stream>>>(d_mons, bt, (long long)PX * PY, PX);
      ++launches;
      bt.n = 0; used = 0; maxcells = 0;
    };
    for (size_t q = 0; q < monitors.size(); ++q) {
      Monitor& m = monitors[q];
      if ((m.comp >= 3) != (group == 0)) continue;
      if (!m.local) continue;
      if (m.decimation > 1 && (timestep % m.decimation) != 0) continue;  // Kernels.jl:465-497
      if (!norm_stale.empty()) norm_stale[q] = 1;
      int nf = (int)m.freqs.size();
      for (int k0 = 0; k0 < nf;) {
        if (bt.n == DFT_SLOTS || used == DFT_PHASORS) flush();
        int kc = std::min(nf - k0, DFT_PHASORS - used);
        int sl = bt.n++;
        bt.mon[sl] = (int)q; bt.k0[sl] = k0; bt.kc[sl] = kc; bt.off[sl] = used;
        for (int k = 0; k < kc; ++k) {
          // phase = T(f_k) * T(2 pi t) formed in T; exp(i phase) = (cos, sin) rounded to T
          T ph = m.freqs[k0 + k] * tf;
          T er = (T)std::cos((double)ph), ei = (T)std::sin((double)ph);
          bt.ph_re[used + k] = dt * er;
          bt.ph_im[used + k] = dt * ei;
        }
        used += kc;
        k0 += kc;
        maxcells = std::max(maxcells, (long long)m.n[0] * m.n[1] * std::max(0, monLocalNz(q)));
      }
    }
    flush();
    CUDA_OK(cudaGetLastError());
  }
  int monLocalNz(size_t q) const { return mon_local_nz[q]; }
  void step(int n) override {
    need_final();
    launches = 0;
    CUDA_OK(cudaEventRecord(ev_t0, stream));
    for (int i = 0; i < n; ++i) {
      double t = time_now();
      double th = t + (double)(dt / T(2));
      update_sources_active(t);
      if (sweep) {
        // the H fields are not touched by the E update, so their DFT may follow the whole step
        sweep_step(t, th);
        dft_update(0, t);
        dft_update(1, th);
      } else if (tma_fuse) {
        fused_step(t, th);
        dft_update(0, t);
        dft_update(1, th);
      } else if (graph_usable() && plain_steps_done >= 1) {
        // H is not touched by the E half-step, so its DFT may follow the whole step
        graph_step(t, th);
        dft_update(0, t);
        dft_update(1, th);
      } else {
        half_step_all(0, t);
        dft_update(0, t);
        half_step_all(1, th);
        dft_update(1, th);
        plain_steps_done += 1;
      }
      timestep += 1;
      if (im) im->timestep = timestep;
    }
    CUDA_OK(cudaEventRecord(ev_t1, stream));
    last_launches = launches;
    last_ms = -1;
  }
  void reset_fields() override {
    for (int c = 0; c < 6; ++c) CUDA_OK(cudaMemsetAsync(F[c], 0, fsize * sizeof(T), stream));
    for (int gq = 0; gq < 2; ++gq)
      for (int d = 0; d < 3; ++d) {
        int nx = (d + 1) % 3;
        if (W[gq][d]) CUDA_OK(cudaMemsetAsync(W[gq][d], 0, slab_elems[d] * sizeof(T), stream));
        if (U[gq][d]) CUDA_OK(cudaMemsetAsync(U[gq][d], 0, slab_elems[nx] * sizeof(T), stream));
        if (Cst[gq][d]) CUDA_OK(cudaMemsetAsync(Cst[gq][d], 0, msize * sizeof(T), stream));
      }
    for (auto& p : poles)
      for (int q = 0; q < 2; ++q)
        for (int d = 0; d < 3; ++d) CUDA_OK(cudaMemsetAsync(p.P[q][d], 0, msize * sizeof(T), stream));
    for (int d = 0; d < 3; ++d)
      if (Dst[d]) CUDA_OK(cudaMemsetAsync(Dst[d], 0, msize * sizeof(T), stream));
    for (int gq = 0; gq < 2; ++gq)
      if (Tsrc[gq]) CUDA_OK(cudaMemsetAsync(Tsrc[gq], 0, nslots[gq] * sizeof(T), stream));
    for (auto& m : monitors) CUDA_OK(cudaMemsetAsync(m.M, 0, 2 * m.elems * sizeof(T), stream));
    for (auto& c : norm_stale) c = 1;
    for (auto& s : sources) { s.ao_re = 0; s.ao_im = 0; }
    for (int gq = 0; gq < 2; ++gq) {
      if (d_done[gq]) CUDA_OK(cudaMemsetAsync(d_done[gq], 0, (size_t)nchunk * sizeof(unsigned long long), stream));
      epochs[gq] = 0;
    }
    if (d_fuse_done) CUDA_OK(cudaMemsetAsync(d_fuse_done, 0, (size_t)nchunk * 8, stream));
    fuse_epoch = 0;
    timestep = 0;
    sources_active = true;
  }
  void comm_init(const void* id, int nranks, int rank) override {
    if (nranks != g.nranks || rank != g.rank) throw std::string("khr_comm_init: rank/nranks differ from the grid descriptor");
    load_nccl();
    Id128 u;
    memcpy(u.b, id, 128);
    CUDA_OK(cudaSetDevice(device));
    NCCL_OK(g_nccl.CommInitRank(&comm, nranks, u, rank));
    comm_handle = comm;
    // which monitor boxes are cut by a slab boundary?  A rank that holds a strict, non-empty part of a box knows
    // it; one max-all-reduce tells everybody.  Surface integrals over uncut boxes then need no array reduction:
    // the owner's accumulators are the whole box, everybody else's are zero (khr_flux: sum of nf doubles).
    mon_split.assign(monitors.size(), 0);
    if (!monitors.empty() && finalized) {
      std::vector<int> h(monitors.size(), 0);
      for (size_t q = 0; q < monitors.size(); ++q)
        h[q] = (mon_local_nz[q] > 0 && mon_local_nz[q] < monitors[q].n[2]) ? 1 : 0;
      int* d = (int*)dalloc((sizeof(int) * h.size() + sizeof(T) - 1) / sizeof(T) + 4, false);
      CUDA_OK(cudaMemcpyAsync(d, h.data(), sizeof(int) * h.size(), cudaMemcpyHostToDevice, comm_stream));
      NCCL_OK(g_nccl.AllReduce(d, d, h.size(), /*ncclInt32*/ 2, /*ncclMax*/ 2, comm, comm_stream));
      CUDA_OK(cudaMemcpyAsync(h.data(), d, sizeof(int) * h.size(), cudaMemcpyDeviceToHost, comm_stream));
      CUDA_OK(cudaStreamSynchronize(comm_stream));
      mon_split = h;
      mon_split_known = true;
    }
  }
  // complex fields: the imaginary parts exchange their halo planes on a communicator of their own (same ranks,
  // ncclCommSplit of the real context's), so the two exchanges of a half-step overlap instead of serialising
  void comm_split_from(Base* parent) override {
    if (!parent->comm_handle) throw std::string("khr_comm_init: the real context has no communicator");
    if (!g_nccl.CommSplit) throw std::string("complex fields on several ranks need ncclCommSplit (NCCL >= 2.18)");
    CUDA_OK(cudaSetDevice(device));
    NCCL_OK(g_nccl.CommSplit(parent->comm_handle, 0, g.rank, &comm, nullptr));
    comm_handle = comm;
  }
  std::vector<int> mon_split;
  bool mon_split_known = false;
  void halo_exchange(int gq) override {
    post_halo(gq);
    wait_halo();
    CUDA_OK(cudaStreamSynchronize(stream));
  }

  void field_read(int comp, void* out) override {
    std::vector<T> h(fsize);
    CUDA_OK(cudaStreamSynchronize(stream));
    CUDA_OK(cudaMemcpy(h.data(), F[comp], fsize * sizeof(T), cudaMemcpyDeviceToHost));
    T* o = (T*)out;
    for (int z = 1; z <= N[2]; ++z)
      for (int y = 1; y <= N[1]; ++y)
        memcpy(o + (size_t)N[0] * ((size_t)(y - 1) + (size_t)N[1] * (z - 1)), h.data() + fidx(1, y, z), N[0] * sizeof(T));
  }
  void field_write(int comp, const void* in) override {
    need_final();
    if (comp < 3 && (!poles.empty() || chi3))
      throw std::string("khr_field_write: writing E is not supported when dispersive poles or a Kerr coefficient "
                        "are registered (the D / P history would be inconsistent); use khr_reset_fields");
    for (auto& s : sources)
      if ((s.comp >= 3) == (comp >= 3))
        throw std::string("khr_field_write: not supported for a field group that has sources registered "
                          "(the flux accumulators of the source voxels would be inconsistent)");
    std::vector<T> h(fsize, T(0));
    const T* src = (const T*)in;
    for (int z = 1; z <= N[2]; ++z)
      for (int y = 1; y <= N[1]; ++y)
        memcpy(h.data() + fidx(1, y, z), src + (size_t)N[0] * ((size_t)(y - 1) + (size_t)N[1] * (z - 1)), N[0] * sizeof(T));
    CUDA_OK(cudaStreamSynchronize(stream));
    CUDA_OK(cudaMemcpy(F[comp], h.data(), fsize * sizeof(T), cudaMemcpyHostToDevice));
    // keep the W history consistent with the stored field (W = m^-1 * net = A)
    int gq = comp >= 3 ? 0 : 1, d = comp % 3;
    if (W[gq][d]) {
      std::vector<T> w(slab_elems[d] + 64, T(0));
      for (int z = 1; z <= N[2]; ++z)
        for (int y = 1; y <= N[1]; ++y)
          for (int x = 1; x <= N[0]; ++x) {
            int i[3] = {x, y, z};
            if (h_sig[gq][d].empty() || h_sig[gq][d][i[d] - 1] == T(0)) continue;
            size_t wi;
            if (d == 0) wi = (size_t)(slab[0].idx(1 + 4 * ((x - 1) / 4)) + (x - 1) % 4) + (size_t)cxp * ((size_t)(y - 1) + (size_t)N[1] * (z - 1));
            else if (d == 1) wi = (size_t)(x - 1) + (size_t)MPX * ((size_t)slab[1].idx(y) + (size_t)cy * (z - 1));
            else wi = (size_t)(x - 1) + (size_t)MPX * ((size_t)(y - 1) + (size_t)N[1] * slab[2].idx(z));
            w[wi] = h[fidx(x, y, z)];
          }
      CUDA_OK(cudaMemcpy(W[gq][d], w.data(), slab_elems[d] * sizeof(T), cudaMemcpyHostToDevice));
    }
  }
  void field_view(int comp, void** p, int64_t* stride, int64_t* off) override {
    *p = F[comp];
    stride[0] = 1; stride[1] = PX; stride[2] = (int64_t)PX * PY;
    *off = XO;
  }
  void monitor_read(int id, void* out) override {
    if (id < 0 || id >= (int)monitors.size()) throw std::string("bad monitor id");
    CUDA_OK(cudaStreamSynchronize(stream));
    CUDA_OK(cudaMemcpy(out, monitors[id].M, 2 * monitors[id].elems * sizeof(T), cudaMemcpyDeviceToHost));
  }
  void monitor_view(int id, void** p, int64_t* dims) override {
    if (id < 0 || id >= (int)monitors.size()) throw std::string("bad monitor id");
    Monitor& m = monitors[id];
    *p = m.M;
    dims[0] = m.n[0]; dims[1] = m.n[1]; dims[2] = m.n[2]; dims[3] = (int64_t)m.freqs.size();
  }
  // Simulation.jl:440-445 stop_when_dft_decayed evaluates every monitor's norm on every step;
  // the accumulators only change on a DFT update, so the values are cached in between and
  // all stale ones are reduced by one launch + one small device->host copy.
  std::vector<double> norm_cache;
  std::vector<char> norm_stale;
  static constexpr int NORM_BLOCKS = 148 * 4;
  double* d_norms = nullptr;      // [n] results, then [n * NORM_BLOCKS] block partials
  int* d_norm_nb = nullptr;
  double* h_norms = nullptr;
  int* h_norm_nb = nullptr;
  void norms_setup() {
    const int n = (int)monitors.size();
    if (d_norms || n == 0) return;
    d_norms = (double*)dalloc((sizeof(double) * (size_t)n * (NORM_BLOCKS + 1) + sizeof(T) - 1) / sizeof(T));
    d_norm_nb = (int*)dalloc((sizeof(int) * (size_t)n + sizeof(T) - 1) / sizeof(T));
    CUDA_OK(cudaMallocHost((void**)&h_norms, sizeof(double) * n));
    CUDA_OK(cudaMallocHost((void**)&h_norm_nb, sizeof(int) * n));
    norm_cache.assign(n, 0.0);
    norm_stale.assign(n, 1);
  }
  // reduce the monitors flagged in `which` (1 = reduce); results land in h_norms (sum of squares)
  void norms_reduce(const std::vector<char>& which) {
    const int n = (int)monitors.size();
    for (int q = 0; q < n; ++q) {
      h_norm_nb[q] = 0;
      if (!which[q]) continue;
      const long long ne = 2 * (long long)monitors[q].elems;
      const int blocks = (int)std::min<long long>((ne + 255) / 256, NORM_BLOCKS);
      h_norm_nb[q] = blocks;
      if (blocks > 0) { sumsq_kernel<T><<<blocks, 256, 0, stream>>>(monitors[q].M, ne, d_norms + n + (size_t)q * NORM_BLOCKS); ++launches; }
    }
    CUDA_OK(cudaMemcpyAsync(d_norm_nb, h_norm_nb, sizeof(int) * n, cudaMemcpyHostToDevice, stream));
    sumsq_finish_kernel<<<(n + 63) / 64, 64, 0, stream>>>(d_norms + n, d_norm_nb, NORM_BLOCKS, n, d_norms);
    ++launches;
    CUDA_OK(cudaGetLastError());
    CUDA_OK(cudaMemcpyAsync(h_norms, d_norms, sizeof(double) * n, cudaMemcpyDeviceToHost, stream));
    CUDA_OK(cudaStreamSynchronize(stream));
  }
  void monitor_norms(double* out, int count) override {
    int n = (int)monitors.size();
    if (count != n) throw std::string("khr_monitor_norms: count must equal the number of monitors");
    if (n == 0) return;
    norms_setup();
    bool any = false;
    for (int q = 0; q < n; ++q) any |= norm_stale[q] != 0;
    if (any) {
      norms_reduce(norm_stale);
      for (int q = 0; q < n; ++q)
        if (norm_stale[q]) { norm_cache[q] = monitors[q].elems ? std::sqrt(h_norms[q]) : 0.0; norm_stale[q] = 0; }
    }
    for (int q = 0; q < n; ++q) out[q] = norm_cache[q];
  }
  // FluxMonitor.jl:92-156 get_flux on the device: only nf doubles cross PCIe instead of the
  // four DFT arrays (Array(md.fields), FluxMonitor.jl:99-102)
  double* d_flux = nullptr;
  size_t d_flux_cap = 0;
  void flux(const int32_t* ids4, int normal_axis, double* out, int nfreq) override {
    bool local_only = false;
    FluxArgs<T> a = surface_args("khr_flux", ids4, normal_axis, nfreq, &local_only);
    const bool empty = a.n1 < 1 || a.n2 < 1;
    if (empty && !local_only) { for (int k = 0; k < nfreq; ++k) out[k] = 0.0; return; }
    const int nblocks = empty ? 1 : (int)std::min<long long>(((long long)a.n1 * a.n2 + 255) / 256, 256);
    double* buf = scratch_f64((size_t)nfreq * (nblocks + 1));
    if (empty) CUDA_OK(cudaMemsetAsync(buf, 0, sizeof(double) * nfreq, stream));
    else {
      flux_kernel<T><<<dim3((unsigned)nblocks, (unsigned)nfreq), 256, 0, stream>>>(a, buf + nfreq);
      flux_finish_kernel<<<(nfreq + 63) / 64, 64, 0, stream>>>(buf + nfreq, nblocks, nfreq, buf);
      CUDA_OK(cudaGetLastError());
      launches += 2;
    }
    if (local_only) {
      // the box lies inside one slab: that rank's result is the flux, everybody else computed 0 from empty accumulators
      CUDA_OK(cudaEventRecord(ev_boundary, stream));
      CUDA_OK(cudaStreamWaitEvent(comm_stream, ev_boundary, 0));
      NCCL_OK(g_nccl.AllReduce(buf, buf, (size_t)nfreq, /*ncclFloat64*/ 8, /*ncclSum*/ 0, comm, comm_stream));
      CUDA_OK(cudaEventRecord(ev_comm, comm_stream));
      CUDA_OK(cudaStreamWaitEvent(stream, ev_comm, 0));
    }
    CUDA_OK(cudaMemcpyAsync(out, buf, sizeof(double) * nfreq, cudaMemcpyDeviceToHost, stream));
    CUDA_OK(cudaStreamSynchronize(stream));
  }
  // Several ranks: every rank accumulates the planes it owns and keeps zeros elsewhere, so the sum
  // over ranks IS the single-domain accumulator, bit for bit (x + 0 = x).  The four arrays are summed
  // with ncclAllReduce into scratch copies (the accumulators themselves keep running) and every rank
  // evaluates the surface integral on the complete arrays — the reference reduces the arrays with
  // MPI.Reduce! before its host loops (Visualization.jl:325-331).  COLLECTIVE: all ranks must call
  // the same surface function for the same monitors in the same order.
  T* surf_scratch[4] = {nullptr, nullptr, nullptr, nullptr};
  size_t surf_cap[4] = {0, 0, 0, 0};
  FluxArgs<T> surface_args(const char* who, const int32_t* ids4, int normal_axis, int nfreq, bool* local_only = nullptr) {
    need_final();
    if (local_only) *local_only = false;
    if (normal_axis < 0 || normal_axis > 2) throw std::string(who) + ": normal axis must be 0, 1 or 2";
    FluxArgs<T> a;
    a.normal = normal_axis;
    a.t1 = normal_axis == 0 ? 1 : 0;
    a.t2 = normal_axis == 2 ? 1 : 2;
    a.n1 = a.n2 = 1 << 30;
    for (int q = 0; q < 4; ++q) {
      if (ids4[q] < 0 || ids4[q] >= (int)monitors.size()) throw std::string(who) + ": bad monitor id";
      const Monitor& m = monitors[ids4[q]];
      if ((int)m.freqs.size() != nfreq) throw std::string(who) + ": the four monitors must share the frequency list";
      if ((q < 2) != (m.comp < 3)) throw std::string(who) + ": monitors must be ordered E1, E2, H1, H2";
      a.M[q] = m.M;
      for (int d = 0; d < 3; ++d) a.n[q][d] = m.n[d];
      if (m.n[normal_axis] < 1) throw std::string(who) + ": empty monitor box";
      a.n1 = std::min(a.n1, m.n[a.t1]);
      a.n2 = std::min(a.n2, m.n[a.t2]);
    }
    a.nf = nfreq;
    a.dA = (double)dl[a.t1] * (double)dl[a.t2];
    if (g.nranks > 1 && local_only && mon_split_known) {
      bool cut = false;
      for (int q = 0; q < 4; ++q) cut = cut || mon_split[(size_t)ids4[q]] != 0;
      if (!cut) { *local_only = true; return a; }    // the caller sums its result over the ranks instead
    }
    if (g.nranks > 1) {
      if (!comm) throw std::string(who) + ": nranks > 1 but khr_comm_init was not called";
      CUDA_OK(cudaEventRecord(ev_boundary, stream));
      CUDA_OK(cudaStreamWaitEvent(comm_stream, ev_boundary, 0));
      for (int q = 0; q < 4; ++q) {
        const Monitor& m = monitors[ids4[q]];
        const size_t cnt = 2 * m.elems;
        if (cnt > surf_cap[q]) { surf_scratch[q] = dalloc(cnt, false); surf_cap[q] = cnt; }
        NCCL_OK(g_nccl.AllReduce(m.M, surf_scratch[q], cnt, sizeof(T) == 4 ? /*ncclFloat32*/ 7 : /*ncclFloat64*/ 8, /*ncclSum*/ 0, comm, comm_stream));
        a.M[q] = surf_scratch[q];
      }
      CUDA_OK(cudaEventRecord(ev_comm, comm_stream));
      CUDA_OK(cudaStreamWaitEvent(stream, ev_comm, 0));
    }
    return a;
  }
  double* scratch_f64(size_t need) {
    if (need > d_flux_cap) {
      d_flux = (double*)dalloc((sizeof(double) * need + sizeof(T) - 1) / sizeof(T), false);
      d_flux_cap = need;
    }
    return d_flux;
  }
  // Near2Far.jl:254-371 / :380-560 compute_far_field at explicit observation points: EH (nobs, 6, nf)
  // ComplexF64.  The DFT arrays stay on the device; obs in, 6 nobs nf complex numbers out.
  void near2far(const int32_t* ids4, int normal_axis, double normal_sign, double eps, double mu, const double* base12,
                const double* freqs, int nfreq, const double* obs, int nobs, double* out) override {
    N2FArgs<T> a;
    a.s = surface_args("khr_near2far", ids4, normal_axis, nfreq);
    if (!(eps > 0) || !(mu > 0)) throw std::string("khr_near2far: medium eps and mu must be positive");
    if (nobs < 1) throw std::string("khr_near2far: no observation points");
    if (nfreq > 65535) throw std::string("khr_near2far: too many frequencies");
    const size_t nout = (size_t)2 * 6 * nobs * nfreq;
    if (a.s.n1 < 1 || a.s.n2 < 1) { for (size_t k = 0; k < nout; ++k) out[k] = 0.0; return; }
    for (int q = 0; q < 4; ++q)
      for (int d = 0; d < 3; ++d) a.base[q][d] = base12[3 * q + d];
    a.d1 = (double)dl[a.s.t1]; a.d2 = (double)dl[a.s.t2];
    a.ns = normal_sign; a.eps = eps; a.mu = mu;
    a.nobs = nobs;
    // enough CTAs to fill the GPU several times over even for a handful of observation points
    const long long ncell = (long long)a.s.n1 * a.s.n2;
    long long nchunk = (148 * 8 + (long long)nobs * nfreq - 1) / ((long long)nobs * nfreq);
    nchunk = std::max<long long>(1, std::min<long long>(nchunk, (ncell + 255) / 256));
    nchunk = std::min<long long>(nchunk, 65535);
    a.nchunk = (int)nchunk;
    const size_t n_part = (size_t)12 * nchunk * nfreq * nobs;
    double* buf = scratch_f64(n_part + nout + (size_t)3 * nobs + nfreq);
    double* d_part = buf;
    double* d_out = d_part + n_part;
    double* d_obs = d_out + nout;
    double* d_fr = d_obs + (size_t)3 * nobs;
    CUDA_OK(cudaMemcpyAsync(d_obs, obs, sizeof(double) * 3 * nobs, cudaMemcpyHostToDevice, stream));
    CUDA_OK(cudaMemcpyAsync(d_fr, freqs, sizeof(double) * nfreq, cudaMemcpyHostToDevice, stream));
    a.obs = d_obs; a.freqs = d_fr;
    near2far_kernel<T><<<dim3((unsigned)nobs, (unsigned)nfreq, (unsigned)nchunk), 256, 0, stream>>>(a, d_part);
    const long long nt = (long long)nobs * nfreq * 6;
    near2far_finish_kernel<<<(unsigned)((nt + 255) / 256), 256, 0, stream>>>(d_part, (int)nchunk, nfreq, nobs, d_out);
    CUDA_OK(cudaGetLastError());
    launches += 2;
    CUDA_OK(cudaMemcpyAsync(out, d_out, sizeof(double) * nout, cudaMemcpyDeviceToHost, stream));
    CUDA_OK(cudaStreamSynchronize(stream));
  }
  // ModeMonitor.jl:462-498: P_mode and the two overlap sums per frequency; out5 = nf x
  // (P, re o+, im o+, re o-, im o-).  mode: ComplexF64 [4][nf][n2][n1] (e1, e2, h1, h2 on the DFT grid)
  void mode_overlap(const int32_t* ids4, int normal_axis, const double* mode, int n1, int n2, int nfreq, double* out5) override {
    FluxArgs<T> a = surface_args("khr_mode_overlap", ids4, normal_axis, nfreq);
    if (n1 != a.n1 || n2 != a.n2)
      throw std::string("khr_mode_overlap: the mode profile extent must equal the common tangential extent of the monitors (") +
          std::to_string(a.n1) + " x " + std::to_string(a.n2) + ")";
    if (nfreq > 65535) throw std::string("khr_mode_overlap: too many frequencies");
    if (a.n1 < 1 || a.n2 < 1) { for (int k = 0; k < 5 * nfreq; ++k) out5[k] = 0.0; return; }
    const size_t ncell = (size_t)n1 * n2;
    const int nblocks = (int)std::min<size_t>((ncell + 255) / 256, 256);
    const size_t n_mode = (size_t)2 * 4 * nfreq * ncell, n_part = (size_t)5 * nfreq * nblocks;
    double* buf = scratch_f64(n_mode + n_part + (size_t)5 * nfreq);
    double* d_mode = buf;
    double* d_part = d_mode + n_mode;
    double* d_out = d_part + n_part;
    CUDA_OK(cudaMemcpyAsync(d_mode, mode, sizeof(double) * n_mode, cudaMemcpyHostToDevice, stream));
    mode_overlap_kernel<T><<<dim3((unsigned)nblocks, (unsigned)nfreq), 256, 0, stream>>>(a, d_mode, d_part);
    mode_overlap_finish_kernel<<<(5 * nfreq + 63) / 64, 64, 0, stream>>>(d_part, nblocks, nfreq, d_out);
    CUDA_OK(cudaGetLastError());
    launches += 2;
    CUDA_OK(cudaMemcpyAsync(out5, d_out, sizeof(double) * 5 * nfreq, cudaMemcpyDeviceToHost, stream));
    CUDA_OK(cudaStreamSynchronize(stream));
  }
  // DiffractionMonitor.jl:87-165: power of the diffraction orders |m|, |n| <= max_order per frequency
  void diffraction(const int32_t* ids4, int normal_axis, int max_order, double L1, double L2, double kinc1, double kinc2,
                   const double* freqs, int nfreq, double* power, int32_t* propagating) override {
    DiffArgs<T> a;
    a.s = surface_args("khr_diffraction", ids4, normal_axis, nfreq);
    if (max_order < 0 || max_order > 64) throw std::string("khr_diffraction: max_order must be 0..64");
    if (!(L1 > 0) || !(L2 > 0)) throw std::string("khr_diffraction: cell sizes must be positive");
    if (nfreq > 65535) throw std::string("khr_diffraction: too many frequencies");
    const int nord = 2 * max_order + 1;
    const size_t nout = (size_t)nfreq * nord * nord;
    if (a.s.n1 < 1 || a.s.n2 < 1) { for (size_t k = 0; k < nout; ++k) { power[k] = 0.0; propagating[k] = 0; } return; }
    a.max_order = max_order; a.L1 = L1; a.L2 = L2; a.kinc1 = kinc1; a.kinc2 = kinc2;
    double* buf = scratch_f64(nout + nout / 2 + 2 + (size_t)nfreq);
    double* d_out = buf;
    int* d_prop = (int*)(d_out + nout);
    double* d_fr = d_out + nout + nout / 2 + 2;
    CUDA_OK(cudaMemcpyAsync(d_fr, freqs, sizeof(double) * nfreq, cudaMemcpyHostToDevice, stream));
    a.freqs = d_fr;
    diffraction_kernel<T><<<dim3((unsigned)(nord * nord), (unsigned)nfreq), 256, 0, stream>>>(a, d_out, d_prop);
    CUDA_OK(cudaGetLastError());
    launches += 1;
    CUDA_OK(cudaMemcpyAsync(power, d_out, sizeof(double) * nout, cudaMemcpyDeviceToHost, stream));
    CUDA_OK(cudaMemcpyAsync(propagating, d_prop, sizeof(int) * nout, cudaMemcpyDeviceToHost, stream));
    CUDA_OK(cudaStreamSynchronize(stream));
  }
  double monitor_norm(int id) override {
    if (id < 0 || id >= (int)monitors.size()) throw std::string("bad monitor id");
    norms_setup();
    std::vector<char> which(monitors.size(), 0);
    which[(size_t)id] = 1;
    norms_reduce(which);
    return monitors[(size_t)id].elems ? std::sqrt(h_norms[id]) : 0.0;
  }
  void sync() override {
    sync_all();
    CUDA_OK(cudaStreamSynchronize(comm_stream));
    if (h_err && *h_err) { *h_err = 0; throw std::string("a step kernel timed out waiting for its H/E dependency counters (chain / sweep / fused mode)"); }
    if (last_ms < 0) {
      float ms = 0;
      if (cudaEventElapsedTime(&ms, ev_t0, ev_t1) == cudaSuccess) last_ms = ms;
      else { cudaGetLastError(); last_ms = 0; }
    }
  }
  void census(int64_t* c) override {
    need_final();
    int64_t np[3], nn[3];
    for (int a = 0; a < 3; ++a) {
      int64_t k = 0;
      for (int i = 1; i <= N[a]; ++i) {
        bool on = false;
        for (int gq = 0; gq < 2; ++gq)
          if (have_sigma[gq][a] && h_sig[gq][a][i - 1] != T(0)) on = true;
        k += on;
      }
      np[a] = k; nn[a] = N[a] - k;
    }
    c[0] = nn[0] * nn[1] * nn[2];
    c[1] = np[0] * nn[1] * nn[2] + nn[0] * np[1] * nn[2] + nn[0] * nn[1] * np[2];
    c[2] = np[0] * np[1] * nn[2] + np[0] * nn[1] * np[2] + nn[0] * np[1] * np[2];
    c[3] = np[0] * np[1] * np[2];
  }
};

}  // namespace khr

// ============================================================================
// C ABI
// ============================================================================
struct khr_ctx {
  khr::Base* impl;
  khr::Base* impl_im = nullptr;   // imaginary parts of the fields (khr_set_complex_fields)
};
// registrations that describe the medium go to both parts of a complex simulation
#define KHR_BOTH(call)                                   \
  do {                                                   \
    ctx->impl->registered_any = true;                    \
    ctx->impl->call;                                     \
    if (ctx->impl_im) ctx->impl_im->call;                \
  } while (0)

#define KHR_TRY(...)                               \
  try {                                            \
    __VA_ARGS__;                                   \
    return 0;                                      \
  } catch (const std::string& e) {                 \
    return khr::fail(e);                           \
  } catch (const std::exception& e) {              \
    return khr::fail(e.what());                    \
  } catch (...) {                                  \
    return khr::fail("unknown error");             \
  }
#define NEED_CTX                                   \
  if (!ctx || !ctx->impl) return khr::fail("null context");

extern "C" {

const char* khr_last_error(void) { return khr::g_err.c_str(); }
int32_t khr_version(void) { return 110; }

int32_t khr_ctx_create(int32_t device, const khr_grid_desc* grid, khr_ctx** out) {
  if (!grid || !out) return khr::fail("null argument");
  KHR_TRY({
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0)
      throw std::string("no CUDA device available (libkhronos_b200 has no CPU fallback): ") + cudaGetErrorString(e);
    if (device < 0 || device >= ndev) throw std::string("bad device index");
    if (grid->n[0] < 1 || grid->n[1] < 1 || grid->n[2] < 1 || grid->nz_local < 1)
      throw std::string("3-D grids only (Nx,Ny,Nz >= 1); 2-D (Nz == 0) is not supported");
    if (grid->z_start < 1 || grid->z_start + grid->nz_local - 1 > grid->n[2]) throw std::string("bad z slab");
    if (grid->nranks < 1 || grid->rank < 0 || grid->rank >= grid->nranks) throw std::string("bad rank");
    khr_ctx* c = new khr_ctx;
    if (grid->dtype == KHR_F32) c->impl = new khr::Impl<float>(device, *grid);
    else if (grid->dtype == KHR_F64) c->impl = new khr::Impl<double>(device, *grid);
    else { delete c; throw std::string("dtype must be KHR_F32 or KHR_F64"); }
    *out = c;
  })
}
int32_t khr_ctx_destroy(khr_ctx* ctx) {
  if (!ctx) return 0;
  KHR_TRY({ delete ctx->impl; delete ctx->impl_im; delete ctx; })
}
int32_t khr_set_pml_sigma(khr_ctx* ctx, int32_t group, int32_t axis, const void* sigma, int32_t len) {
  NEED_CTX
  if (group < 0 || group > 1 || axis < 0 || axis > 2 || !sigma) return khr::fail("bad argument");
  KHR_TRY(KHR_BOTH(set_pml_sigma(group, axis, sigma, len)))
}
int32_t khr_set_grid_spacing(khr_ctx* ctx, int32_t axis, const void* spacing, int32_t len) {
  NEED_CTX
  if (axis < 0 || axis > 2 || !spacing) return khr::fail("bad argument");
  KHR_TRY(KHR_BOTH(set_grid_spacing(axis, spacing, len)))
}
int32_t khr_geometry_rasterize(khr_ctx* ctx, const khr_object* objects, int32_t nobj, int32_t kinds_mask, int32_t smoothing,
                               const double origins[18], int64_t smoothed_out[3]) {
  NEED_CTX
  if (!objects || !origins) return khr::fail("bad argument");
  KHR_TRY({
    ctx->impl->registered_any = true;
    ctx->impl->geometry_rasterize(objects, nobj, kinds_mask, smoothing, origins, smoothed_out);
    if (ctx->impl_im) ctx->impl_im->geometry_rasterize(objects, nobj, kinds_mask, smoothing, origins, nullptr);
  })
}
int32_t khr_material_read(khr_ctx* ctx, int32_t kind, int32_t comp, void* dense_out) {
  NEED_CTX
  if (!dense_out) return khr::fail("null array");
  KHR_TRY(ctx->impl->material_read(kind, comp, dense_out))
}
int32_t khr_set_material_scalar(khr_ctx* ctx, int32_t kind, double value) {
  NEED_CTX
  KHR_TRY(KHR_BOTH(set_material_scalar(kind, value)))
}
int32_t khr_set_material_array(khr_ctx* ctx, int32_t kind, int32_t comp, const void* dense) {
  NEED_CTX
  if (!dense) return khr::fail("null array");
  KHR_TRY(KHR_BOTH(set_material_array(kind, comp, dense)))
}
int32_t khr_pole_register(khr_ctx* ctx, double omega0, double gamma, const void* sigma_dense, int32_t* pole_id) {
  NEED_CTX
  if (!sigma_dense) return khr::fail("null array");
  KHR_TRY({
    ctx->impl->registered_any = true;
    int id = ctx->impl->pole_register(omega0, gamma, sigma_dense);
    if (ctx->impl_im) ctx->impl_im->pole_register(omega0, gamma, sigma_dense);
    if (pole_id) *pole_id = id;
  })
}
int32_t khr_source_register(khr_ctx* ctx, int32_t comp, const int32_t start[3], const int32_t dims[3],
                            const void* amp_complex, int32_t time_kind, const double tp[4], int32_t* source_id) {
  NEED_CTX
  if (comp < 0 || comp > 5 || !start || !dims || !amp_complex || !tp) return khr::fail("bad argument");
  if (dims[0] < 1 || dims[1] < 1 || dims[2] < 1) return khr::fail("empty source box");
  // complex fields: sources drive the real part only (`+= real(a * A)`, Sources.jl:355-356)
  KHR_TRY({ ctx->impl->registered_any = true; int id = ctx->impl->source_register(comp, start, dims, amp_complex, time_kind, tp); if (source_id) *source_id = id; })
}
int32_t khr_source_set_amplitude(khr_ctx* ctx, int32_t source_id, double re, double im) {
  NEED_CTX
  KHR_TRY(ctx->impl->source_set_amplitude(source_id, re, im))
}
int32_t khr_set_sources_active(khr_ctx* ctx, int32_t mode) {
  NEED_CTX
  ctx->impl->sources_mode = mode;
  if (mode == 1) ctx->impl->sources_active = true;
  if (mode == 0) ctx->impl->sources_active = false;
  return 0;
}
int32_t khr_monitor_register(khr_ctx* ctx, int32_t comp, const int32_t start[3], const int32_t end[3], int32_t nfreq,
                             const double* freqs, int32_t decimation, int32_t* monitor_id) {
  NEED_CTX
  if (comp < 0 || comp > 5 || !start || !end || nfreq < 1 || !freqs) return khr::fail("bad argument");
  KHR_TRY({ ctx->impl->registered_any = true; int id = ctx->impl->monitor_register(comp, start, end, nfreq, freqs, decimation); if (monitor_id) *monitor_id = id; })
}
int32_t khr_set_periodic(khr_ctx* ctx, int32_t axis, int32_t on) {
  NEED_CTX
  if (axis < 0 || axis > 2) return khr::fail("axis must be 0, 1 or 2");
  ctx->impl->periodic[axis] = on != 0;
  if (ctx->impl_im) ctx->impl_im->periodic[axis] = on != 0;
  return 0;
}
int32_t khr_set_complex_fields(khr_ctx* ctx) {
  NEED_CTX
  KHR_TRY({
    if (ctx->impl_im) return 0;
    if (ctx->impl->registered_any) throw std::string("khr_set_complex_fields must be called right after khr_ctx_create, before any registration");
    const khr_grid_desc& g = ctx->impl->g;
    if (g.dtype == KHR_F32) ctx->impl_im = new khr::Impl<float>(ctx->impl->device, g);
    else ctx->impl_im = new khr::Impl<double>(ctx->impl->device, g);
    ctx->impl->set_partner(ctx->impl_im);
  })
}
int32_t khr_set_bloch(khr_ctx* ctx, int32_t axis, double k_times_L) {
  NEED_CTX
  if (axis < 0 || axis > 2) return khr::fail("axis must be 0, 1 or 2");
  if (!ctx->impl_im) return khr::fail("khr_set_bloch needs complex fields: call khr_set_complex_fields first");
  ctx->impl->periodic[axis] = true;
  ctx->impl_im->periodic[axis] = true;
  ctx->impl->bloch_kl[axis] = k_times_L;
  return 0;
}
int32_t khr_finalize_plan(khr_ctx* ctx) {
  NEED_CTX
  KHR_TRY({ if (ctx->impl_im) ctx->impl_im->finalize(); ctx->impl->finalize(); })
}
int32_t khr_step(khr_ctx* ctx, int32_t nsteps) {
  NEED_CTX
  KHR_TRY(ctx->impl->step(nsteps))
}
int32_t khr_step_h(khr_ctx* ctx) {
  NEED_CTX
  KHR_TRY(ctx->impl->step_h())
}
int32_t khr_step_e(khr_ctx* ctx) {
  NEED_CTX
  KHR_TRY(ctx->impl->step_e())
}
int32_t khr_dft_update(khr_ctx* ctx, int32_t group, double time) {
  NEED_CTX
  KHR_TRY(ctx->impl->dft_update(group, time))
}
int32_t khr_get_timestep(khr_ctx* ctx, int64_t* timestep) {
  NEED_CTX
  *timestep = ctx->impl->timestep;
  return 0;
}
int32_t khr_set_timestep(khr_ctx* ctx, int64_t timestep) {
  NEED_CTX
  ctx->impl->timestep = timestep;
  if (ctx->impl_im) ctx->impl_im->timestep = timestep;
  return 0;
}
int32_t khr_reset_fields(khr_ctx* ctx) {
  NEED_CTX
  KHR_TRY({ ctx->impl->reset_fields(); if (ctx->impl_im) ctx->impl_im->reset_fields(); })
}
int32_t khr_comm_unique_id(void* out128) {
  KHR_TRY({ khr::load_nccl(); NCCL_OK(khr::g_nccl.GetUniqueId(out128)); })
}
int32_t khr_comm_init(khr_ctx* ctx, const void* unique_id128, int32_t nranks, int32_t rank) {
  NEED_CTX
  KHR_TRY({ ctx->impl->comm_init(unique_id128, nranks, rank); if (ctx->impl_im) ctx->impl_im->comm_split_from(ctx->impl); })
}
int32_t khr_halo_exchange(khr_ctx* ctx, int32_t group) {
  NEED_CTX
  KHR_TRY(ctx->impl->halo_exchange(group))
}
int32_t khr_field_read(khr_ctx* ctx, int32_t comp, void* dense_out) {
  NEED_CTX
  if (comp < 0 || comp > 5 || !dense_out) return khr::fail("bad argument");
  KHR_TRY(ctx->impl->field_read(comp, dense_out))
}
int32_t khr_field_read_imag(khr_ctx* ctx, int32_t comp, void* dense_out) {
  NEED_CTX
  if (comp < 0 || comp > 5 || !dense_out) return khr::fail("bad argument");
  if (!ctx->impl_im) return khr::fail("khr_field_read_imag: the context has real fields");
  KHR_TRY(ctx->impl_im->field_read(comp, dense_out))
}
int32_t khr_field_write(khr_ctx* ctx, int32_t comp, const void* dense_in) {
  NEED_CTX
  if (comp < 0 || comp > 5 || !dense_in) return khr::fail("bad argument");
  if (ctx->impl_im) return khr::fail("khr_field_write: not supported for complex fields (the imaginary part would keep its old state); use khr_reset_fields");
  KHR_TRY(ctx->impl->field_write(comp, dense_in))
}
int32_t khr_field_view(khr_ctx* ctx, int32_t comp, void** dev_ptr, int64_t stride[3], int64_t* offset) {
  NEED_CTX
  if (comp < 0 || comp > 5) return khr::fail("bad component");
  KHR_TRY(ctx->impl->field_view(comp, dev_ptr, stride, offset))
}
int32_t khr_monitor_read(khr_ctx* ctx, int32_t monitor_id, void* complex_out) {
  NEED_CTX
  KHR_TRY(ctx->impl->monitor_read(monitor_id, complex_out))
}
int32_t khr_monitor_view(khr_ctx* ctx, int32_t monitor_id, void** dev_ptr, int64_t dims[4]) {
  NEED_CTX
  KHR_TRY(ctx->impl->monitor_view(monitor_id, dev_ptr, dims))
}
int32_t khr_monitor_norm(khr_ctx* ctx, int32_t monitor_id, double* norm) {
  NEED_CTX
  KHR_TRY(*norm = ctx->impl->monitor_norm(monitor_id))
}
int32_t khr_monitor_norms(khr_ctx* ctx, double* norms, int32_t count) {
  NEED_CTX
  KHR_TRY(ctx->impl->monitor_norms(norms, count))
}
int32_t khr_flux(khr_ctx* ctx, const int32_t monitor_ids[4], int32_t normal_axis, double* flux_out, int32_t nfreq) {
  NEED_CTX
  if (!monitor_ids || !flux_out || nfreq < 1) return khr::fail("bad argument");
  KHR_TRY(ctx->impl->flux(monitor_ids, normal_axis, flux_out, nfreq))
}
int32_t khr_near2far(khr_ctx* ctx, const int32_t monitor_ids[4], int32_t normal_axis, double normal_sign, double medium_eps,
                     double medium_mu, const double base_xyz[12], const double* freqs, int32_t nfreq, const double* obs_xyz,
                     int32_t nobs, double* eh_out) {
  NEED_CTX
  if (!monitor_ids || !base_xyz || !freqs || !obs_xyz || !eh_out || nfreq < 1) return khr::fail("bad argument");
  KHR_TRY(ctx->impl->near2far(monitor_ids, normal_axis, normal_sign, medium_eps, medium_mu, base_xyz, freqs, nfreq, obs_xyz, nobs, eh_out))
}
int32_t khr_mode_overlap(khr_ctx* ctx, const int32_t monitor_ids[4], int32_t normal_axis, const double* mode_fields, int32_t n1,
                         int32_t n2, int32_t nfreq, double* out5) {
  NEED_CTX
  if (!monitor_ids || !mode_fields || !out5 || nfreq < 1) return khr::fail("bad argument");
  KHR_TRY(ctx->impl->mode_overlap(monitor_ids, normal_axis, mode_fields, n1, n2, nfreq, out5))
}
int32_t khr_diffraction(khr_ctx* ctx, const int32_t monitor_ids[4], int32_t normal_axis, int32_t max_order, double L1, double L2,
                        double kinc1, double kinc2, const double* freqs, int32_t nfreq, double* power, int32_t* propagating) {
  NEED_CTX
  if (!monitor_ids || !freqs || !power || !propagating || nfreq < 1) return khr::fail("bad argument");
  KHR_TRY(ctx->impl->diffraction(monitor_ids, normal_axis, max_order, L1, L2, kinc1, kinc2, freqs, nfreq, power, propagating))
}
int32_t khr_sync(khr_ctx* ctx) {
  NEED_CTX
  KHR_TRY({ ctx->impl->sync(); if (ctx->impl_im) ctx->impl_im->sync(); })
}
int32_t khr_get_stream(khr_ctx* ctx, void** cuda_stream) {
  NEED_CTX
  *cuda_stream = (void*)ctx->impl->stream;
  return 0;
}
int32_t khr_last_step_timing(khr_ctx* ctx, double* ms, int64_t* kernel_launches) {
  NEED_CTX
  KHR_TRY({ ctx->impl->sync(); if (ctx->impl_im) ctx->impl_im->sync(); if (ms) *ms = ctx->impl->last_ms; if (kernel_launches) *kernel_launches = ctx->impl->last_launches; })
}
int32_t khr_voxel_census(khr_ctx* ctx, int64_t counts[4]) {
  NEED_CTX
  KHR_TRY(ctx->impl->census(counts))
}
int32_t khr_set_profiling(khr_ctx* ctx, int32_t mode) {
  NEED_CTX
  KHR_TRY({ ctx->impl->set_profiling(mode); if (ctx->impl_im) ctx->impl_im->set_profiling(mode); })
}
int32_t khr_kernel_stat_get(khr_ctx* ctx, int32_t index, khr_kernel_stat* out, int32_t* count) {
  NEED_CTX
  KHR_TRY({ int n = ctx->impl->kernel_stat(index, out); if (count) *count = n; })
}
int32_t khr_comm_stat_get(khr_ctx* ctx, double* wait_ms, int64_t* exchanges) {
  NEED_CTX
  KHR_TRY(ctx->impl->comm_stat(wait_ms, exchanges))
}
int32_t khr_graph_info(khr_ctx* ctx, int64_t* kernels_per_graph, int64_t* replays) {
  NEED_CTX
  KHR_TRY(ctx->impl->graph_info(kernels_per_graph, replays))
}
int32_t khr_device_bytes(khr_ctx* ctx, int64_t* bytes) {
  NEED_CTX
  *bytes = ctx->impl->dev_bytes + (ctx->impl_im ? ctx->impl_im->dev_bytes : 0);
  return 0;
}

}  // extern "C"
