/*
 * khronos_b200.h — C ABI of libkhronos_b200.so
 *
 * B200-native (sm_100a) replacement of the Khronos.jl FDTD time-step hot path.
 * The reference (pure Julia, /root/reference/src) has no FFI today; every entry
 * point below replaces one Julia call site, cited as file:line relative to the
 * reference tree.  The Julia shim that binds these with `ccall` is shown in
 * INTEGRATION.md; the same ABI is driven from Python (ctypes) by the host-side
 * mirror in khronos.jl_b200/.
 *
 * Conventions
 *   - every function returns int32 status: 0 = ok, non-zero = error; the message
 *     is available from khr_last_error() (thread-local, valid until the next call)
 *   - plain-old-data only: pointers, sizes, scalars.  No callbacks, no exceptions
 *     cross the boundary.
 *   - "dense" host arrays are column-major (x fastest) like Julia arrays, element
 *     type = the context dtype (float or double), extent (Nx,Ny,Nz_local) over the
 *     cells 1..N this context owns, no ghost layers.
 *   - indices are 1-based cell indices exactly as the reference computes them
 *     (GridVolume.start_idx etc.), global in x/y and global in z (the library
 *     subtracts the slab origin itself).
 *   - ownership: the library owns all device storage (padded, 128-byte aligned
 *     rows; SURVEY.md §8(b) convention B).  khr_field_view exposes the raw device
 *     pointer + strides so the host can wrap it zero-copy.
 *   - threading: one caller thread at a time per context; all calls are
 *     asynchronous on the context's stream except khr_sync and the *_read calls.
 */
#ifndef KHRONOS_B200_H
#define KHRONOS_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct khr_ctx khr_ctx;

enum { KHR_F32 = 0, KHR_F64 = 1 };
/* field components, the order of src/Fields.jl / DataStructures.jl:100-148 */
enum { KHR_EX = 0, KHR_EY = 1, KHR_EZ = 2, KHR_HX = 3, KHR_HY = 4, KHR_HZ = 5 };
/* field groups: the reference's :H half-step (B/H from curl E) and :E half-step */
enum { KHR_GROUP_H = 0, KHR_GROUP_E = 1 };
/* per-voxel material arrays (Geometry.jl:460-477).  KHR_MAT_CHI3: the Kerr coefficient of
 * Geometry.jl:610-660 (one array on the centre grid, component 0), consumed by the correction of
 * Dispersive.jl:127-173 that step! applies between the E update and the ADE update */
enum { KHR_MAT_EPS_INV = 0, KHR_MAT_MU_INV = 1, KHR_MAT_SIGMA_D = 2, KHR_MAT_SIGMA_B = 3, KHR_MAT_CHI3 = 4 };
/* time profiles (Sources/TimeSources.jl:61-64, 123-132); HOST = amplitude pushed
 * every step with khr_source_set_amplitude (CustomSourceData, :165-171) */
enum { KHR_TIME_CW = 0, KHR_TIME_GAUSSIAN = 1, KHR_TIME_HOST = 2 };

/* Grid + slab description.  Replaces the SimulationData fields the kernels read
 * (DataStructures.jl:732-741) and, for nranks > 1, the rank's band of the
 * z-slab decomposition (Chunking.jl:696-716, Distributed.jl:104-148). */
typedef struct khr_grid_desc {
  int32_t dtype;        /* KHR_F32 | KHR_F64 */
  int32_t n[3];         /* global cell counts Nx,Ny,Nz */
  double dl[3];         /* Δx,Δy,Δz already rounded to dtype */
  double dt;            /* Δt already rounded to dtype */
  int32_t z_start;      /* first global z cell owned by this context (1-based) */
  int32_t nz_local;     /* number of z cells owned */
  int32_t rank, nranks; /* position in the slab decomposition */
} khr_grid_desc;

const char* khr_last_error(void);
int32_t khr_version(void);

/* Simulation.jl:249-280 (kernel/constant caching block) -> library init */
int32_t khr_ctx_create(int32_t device, const khr_grid_desc* grid, khr_ctx** out);
int32_t khr_ctx_destroy(khr_ctx* ctx);

/* Boundaries.jl:99-164 init_boundaries: upload one 1-D PML profile of length
 * 2N+1 (host, dtype elements); kernels sample sigma[2i-1] (Helpers.jl:277).
 * group: KHR_GROUP_H -> σB*, KHR_GROUP_E -> σD*. */
int32_t khr_set_pml_sigma(khr_ctx* ctx, int32_t group, int32_t axis, const void* sigma, int32_t len);

/* DataStructures.jl:737-739: non-uniform grid — Δ of `axis` as a vector with one spacing per
 * cell (N global values, context dtype).  The curl then multiplies by inv(Δ[i]) of the updated
 * cell (get_inv_dx(Δ::AbstractVector, i), Helpers.jl:283-291); khr_grid_desc.dl keeps the
 * representative scalar the reference uses elsewhere (Δ[1], utils.jl:4-5) and dt =
 * min over all spacings x Courant (DataStructures.jl:692,740).  Before khr_finalize_plan. */
int32_t khr_set_grid_spacing(khr_ctx* ctx, int32_t axis, const void* spacing, int32_t len);

/* Geometry.jl:450-663 init_geometry outputs: scalar ε⁻¹/μ⁻¹ or per-voxel arrays,
 * and the material conductivities σD/σB (absorbers, Geometry.jl:708-789). */
int32_t khr_set_material_scalar(khr_ctx* ctx, int32_t kind, double value);
int32_t khr_set_material_array(khr_ctx* ctx, int32_t kind, int32_t comp, const void* dense);

/* Geometry on the device: replaces the rasterisation loop of init_geometry
 * (Geometry.jl:450-605, _rasterize_object_yrange! :150-246: objects painted last to first inside
 * their bounding-box index ranges, earlier objects win) and _apply_subpixel_smoothing!
 * (:795-972) for scenes of spheres, cuboids and cylinders, writing eps^-1 / mu^-1 / sigma_D / sigma_B
 * straight into the library's device arrays.  Shape predicates follow GeometryPrimitives.jl
 * (`in`, `bounds`, `surfpt_nearby`, `level`, `volfrac`; not vendored by the reference).
 *   objects      : priority order (index 0 wins), painted values already 1/eps etc. in dtype precision
 *   kinds_mask   : bit KHR_MAT_* set for every array to produce (needs_perm / needs_conductivities)
 *   smoothing    : 0 NoSmoothing, 1 VolumeAveraging, 2 AnisotropicSmoothing (DataStructures.jl:66-76)
 *   origins      : get_component_origin (utils.jl:156-170) of Ex,Ey,Ez,Hx,Hy,Hz, 6 x (x,y,z)
 *   smoothed_out : interface voxels rewritten per E component (may be NULL)
 * Uniform grids only; before khr_finalize_plan. */
enum { KHR_SHAPE_SPHERE = 0, KHR_SHAPE_CUBOID = 1, KHR_SHAPE_CYLINDER = 2 };
typedef struct khr_object {
  int32_t kind;
  int32_t pad_;
  double center[3];
  double size[3];     /* sphere: size[0] = radius; cuboid: full edge lengths along its axes; cylinder: radius, height */
  double axes[9];     /* cuboid: rows = axis vectors (orthogonal; normalised by the library); all zero = identity;
                         cylinder: axes[0..2] = axis direction */
  double eps_inv[3], mu_inv[3], sigma_d[3], sigma_b[3];
} khr_object;
int32_t khr_geometry_rasterize(khr_ctx* ctx, const khr_object* objects, int32_t nobj, int32_t kinds_mask, int32_t smoothing,
                               const double origins[18], int64_t smoothed_out[3]);
/* dense (Nx,Ny,Nz_local) copy of a per-voxel material array (kind = KHR_MAT_*) */
int32_t khr_material_read(khr_ctx* ctx, int32_t kind, int32_t comp, void* dense_out);

/* Geometry.jl:1136-1355 init_polarization! output for one pole: σ array shared by
 * x/y/z (Geometry.jl:1291-1302) and the pole parameters; coefficients follow
 * Susceptibility.jl:74-85. */
int32_t khr_pole_register(khr_ctx* ctx, double omega0, double gamma, const void* sigma_dense, int32_t* pole_id);

/* Sources.jl:35-38 add_sources / SourceData: spatial amplitude box (complex,
 * interleaved re/im, extent dims, column-major) at start_idx; time profile
 * parameters tp = {fcen, width, peak_time, cutoff} (dtype-rounded by the caller). */
int32_t khr_source_register(khr_ctx* ctx, int32_t comp, const int32_t start[3], const int32_t dims[3],
                            const void* amp_complex, int32_t time_kind, const double tp[4], int32_t* source_id);
/* Sources.jl:288-328 step_source_chunk!: scalar_amplitude for a KHR_TIME_HOST source */
int32_t khr_source_set_amplitude(khr_ctx* ctx, int32_t source_id, double re, double im);
/* Kernels.jl:27-35: sources_active switch (1 on, 0 off; default: automatic from cutoffs) */
int32_t khr_set_sources_active(khr_ctx* ctx, int32_t mode /* -1 auto, 0 off, 1 on */);

/* Monitors.jl:202-272 init_monitors for one DFTMonitorData: component, index box
 * [start,end] in the component grid, frequencies (dtype-rounded), decimation. */
int32_t khr_monitor_register(khr_ctx* ctx, int32_t comp, const int32_t start[3], const int32_t end[3], int32_t nfreq,
                             const double* freqs, int32_t decimation, int32_t* monitor_id);

/* Chunking.jl:1725-1770 _add_periodic_connections!: both sides of `axis` are Periodic (or
 * Bloch with k = 0): after every half-step the last interior layer is copied to the lower
 * ghost and the first interior layer to the upper ghost, for the three components of the
 * group (z on several ranks: the wrap closes the halo ring).  The caller passes PML
 * thickness 0 on such an axis (eff_boundaries, Boundaries.jl:100-110).  Before
 * khr_finalize_plan.  For Bloch boundaries see khr_set_complex_fields / khr_set_bloch. */
int32_t khr_set_periodic(khr_ctx* ctx, int32_t axis, int32_t on);

/* Fields.jl:140-159 _needs_complex_fields: any Bloch boundary makes every field array Complex{T}.
 * Call right after khr_ctx_create (before any registration).  The library keeps the real and the
 * imaginary parts as two real field sets with identical layout and runs the same real kernels on
 * both (every update coefficient is real, so real x complex products never mix the parts); they
 * meet only in the Bloch phase of the wrap-around copy and in the DFT phasor product.  Sources
 * drive the real part only, as in the reference (`+= real(a(t) * A[x])`, Sources.jl:355-356).
 * Single GPU (nranks == 1), no chi3. */
int32_t khr_set_complex_fields(khr_ctx* ctx);
/* Chunking.jl:1735-1764 Bloch wrap-around on `axis` (both sides Bloch / Periodic): like
 * khr_set_periodic, with the ghost copies multiplied by exp(-i k L) (lower ghost) and exp(+i k L)
 * (upper ghost) in ComplexF64 (:2163-2167).  k_times_L = bloch.k * sim.cell_size[axis] (Float64). */
int32_t khr_set_bloch(khr_ctx* ctx, int32_t axis, double k_times_L);

/* builds region/work tables, allocates PML auxiliary slabs; call once after all
 * registrations (tail of prepare_simulation!, Simulation.jl:198-280) */
int32_t khr_finalize_plan(khr_ctx* ctx);

/* --- stepping ------------------------------------------------------------- */
/* Kernels.jl:20-88 step!: n full time steps (sources, H, H-DFT, E + ADE, E-DFT) */
int32_t khr_step(khr_ctx* ctx, int32_t nsteps);
/* Kernels.jl:164-296 step_H_fused! (+ Sources.jl:330-340 magnetic sources, + halo) */
int32_t khr_step_h(khr_ctx* ctx);
/* Kernels.jl:298-454 step_E_fused! + Dispersive.jl:186-228 step_polarization! (fused) */
int32_t khr_step_e(khr_ctx* ctx);
/* Monitors.jl:274-329 update_monitor for every DFT monitor of the group at `time` */
int32_t khr_dft_update(khr_ctx* ctx, int32_t group, double time);
/* Kernels.jl:87 increment_timestep! / Simulation.jl:22 */
int32_t khr_get_timestep(khr_ctx* ctx, int64_t* timestep);
int32_t khr_set_timestep(khr_ctx* ctx, int64_t timestep);
/* Simulation.jl:571-637 reset_fields! */
int32_t khr_reset_fields(khr_ctx* ctx);

/* --- distributed (Distributed.jl:57-72 init_nccl!, :448-506 halo) --------- */
int32_t khr_comm_unique_id(void* out128);
int32_t khr_comm_init(khr_ctx* ctx, const void* unique_id128, int32_t nranks, int32_t rank);
/* Chunking.jl:2106-2130 exchange_halos!(sim, :H | :E) — blocking form for API parity */
int32_t khr_halo_exchange(khr_ctx* ctx, int32_t group);

/* --- data access ---------------------------------------------------------- */
/* Visualization.jl:294-333 _pull_fields_from_device: dense (Nx,Ny,Nz_local) copy of cells 1..N */
int32_t khr_field_read(khr_ctx* ctx, int32_t comp, void* dense_out);
/* imaginary part of a complex field (khr_set_complex_fields), same layout as khr_field_read */
int32_t khr_field_read_imag(khr_ctx* ctx, int32_t comp, void* dense_out);
int32_t khr_field_write(khr_ctx* ctx, int32_t comp, const void* dense_in);
/* raw device view: element (ix,iy,iz_local) (1-based cells, 0 = lower ghost) lives at
 * ptr[offset + ix*stride[0] + iy*stride[1] + iz*stride[2]] */
int32_t khr_field_view(khr_ctx* ctx, int32_t comp, void** dev_ptr, int64_t stride[3], int64_t* offset);
/* FluxMonitor.jl:99-102 Array(md.fields): complex (nx,ny,nz,nf) interleaved re/im */
int32_t khr_monitor_read(khr_ctx* ctx, int32_t monitor_id, void* complex_out);
int32_t khr_monitor_view(khr_ctx* ctx, int32_t monitor_id, void** dev_ptr, int64_t dims[4]);
/* Simulation.jl:440-445 stop_when_dft_decayed: sqrt(sum |M|^2) of one monitor */
int32_t khr_monitor_norm(khr_ctx* ctx, int32_t monitor_id, double* norm);
/* all monitors at once (count = number registered); values are cached between DFT updates */
int32_t khr_monitor_norms(khr_ctx* ctx, double* norms, int32_t count);

/* FluxMonitor.jl:92-156 get_flux(md): Poynting flux through the plane of four DFT monitors
 * (ids ordered E1, E2, H1, H2 as init_flux_monitor creates them, :18-70) per frequency, reduced
 * on the device with the reference's per-cell arithmetic (two-plane average and products in
 * Complex{T}, area factor and sum in Float64); replaces Array(md.fields) x4 + the host loop.
 * normal_axis 0..2.  Fails if the box is split across ranks. */
int32_t khr_flux(khr_ctx* ctx, const int32_t monitor_ids[4], int32_t normal_axis, double* flux_out, int32_t nfreq);

/* Near2Far.jl:998-1031 compute_far_field at explicit observation points (:254-371 the host loop,
 * :103-247 / :380-560 the KernelAbstractions kernels): surface-equivalence currents J = n x H,
 * M = -n x E of the plane's four DFT monitors (ids ordered E1, E2, H1, H2, init_near2far_monitor
 * :1137-1200) radiated with the full dyadic Green's function green3d! (:40-96).  All arithmetic
 * in Float64 / ComplexF64 like the reference.  base_xyz: physical position of dft[1,1,1] of each
 * monitor (md.e1_base, e2_base, h1_base, h2_base; Monitors.jl:398-420); freqs: md.frequencies;
 * obs_xyz: nobs points (x,y,z); eh_out: ComplexF64 (nobs, 6, nfreq) column-major, interleaved. */
int32_t khr_near2far(khr_ctx* ctx, const int32_t monitor_ids[4], int32_t normal_axis, double normal_sign, double medium_eps,
                     double medium_mu, const double base_xyz[12], const double* freqs, int32_t nfreq, const double* obs_xyz,
                     int32_t nobs, double* eh_out);

/* ModeMonitor.jl:345-515 compute_mode_amplitudes, the two surface sums (:462-498): mode_fields =
 * the mode profile already interpolated onto the DFT grid (:427-457), ComplexF64 [4][nfreq][n2][n1]
 * (e1, e2, h1, h2; n1 fastest), n1 x n2 = the common tangential extent of the four monitors.
 * out5: per frequency P_mode, re/im of overlap_plus, re/im of overlap_minus; the caller forms
 * a± = overlap± / (4 P_mode) (:500-503). */
int32_t khr_mode_overlap(khr_ctx* ctx, const int32_t monitor_ids[4], int32_t normal_axis, const double* mode_fields, int32_t n1,
                         int32_t n2, int32_t nfreq, double* out5);

/* DiffractionMonitor.jl:87-165 get_diffraction_efficiencies: Poynting power of the diffraction
 * orders |m|, |n| <= max_order per frequency from the spatial DFT of the four tangential DFT
 * monitors of a plane (ids ordered E1, E2, H1, H2).  Only the (2 max_order + 1)^2 needed bins are
 * formed (the reference computes the full O(N^2) transform, fft2_manual :185-197).  L1, L2 =
 * md.cell_size along the two tangential axes, kinc1/2 = the tangential components of k_inc.
 * power / propagating: [nfreq][2M+1][2M+1] (n fastest, m = -M..M); evanescent orders, which the
 * reference skips, have propagating = 0 and power = 0.  The common tangential extent of the four
 * monitors is transformed (the reference mixes per-field extents, :126-140). */
int32_t khr_diffraction(khr_ctx* ctx, const int32_t monitor_ids[4], int32_t normal_axis, int32_t max_order, double L1, double L2,
                        double kinc1, double kinc2, const double* freqs, int32_t nfreq, double* power, int32_t* propagating);

int32_t khr_sync(khr_ctx* ctx);
int32_t khr_get_stream(khr_ctx* ctx, void** cuda_stream);

/* --- measurement helpers (bench.py / profiles) ---------------------------- */
/* device time of the last khr_step() call in ms (CUDA events on the ctx stream)
 * and the number of kernels launched by it */
int32_t khr_last_step_timing(khr_ctx* ctx, double* ms, int64_t* kernel_launches);
/* voxel-class census used by the bytes model: counts of cells with 0,1,2,3 PML axes */
int32_t khr_voxel_census(khr_ctx* ctx, int64_t counts[4]);
/* per-kernel live timing: mode 1 = record CUDA events around every step-kernel launch
 * on the ctx stream, 0 = stop, 2 = start and reset the accumulators, 3 = like 2 but with the
 * kernels of a half-step serialised on one stream, so that each event pair times its kernel alone */
typedef struct khr_kernel_stat {
  char name[96];
  int64_t launches;             /* launches timed since the last reset */
  double total_ms;              /* summed CUDA-event time of those launches */
  int64_t cells_per_launch;     /* voxels one launch updates */
  double alg_bytes_per_launch;  /* compulsory bytes of THIS implementation for those voxels (one half-step):
                                 * 9w per voxel (+3w per-voxel material unless the tile is constant),
                                 * +4w per PML axis of the voxel, + conductivity / ADE / Kerr arrays where present */
  int64_t ctas;                 /* thread blocks per launch */
  int64_t uniform_ctas;         /* of those, tiles with constant per-voxel material (loads skipped) */
  double ref_model_bytes_per_launch; /* the reference's own byte model for the same voxels (SURVEY.md §8(d):
                                 * 9w/12w + 10w/14w/18w on 1/2/3-PML-axis voxels): "bytes saved vs reference" */
} khr_kernel_stat;
/* multi-GPU: time the main stream spent waiting for the halo receive since the last profiling reset
 * (CUDA events around the wait; only recorded while profiling is on), and the number of exchanges */
int32_t khr_comm_stat_get(khr_ctx* ctx, double* wait_ms, int64_t* exchanges);
/* khr_step can replay a CUDA graph of one time step (single GPU; the reference captures graphs of its
 * step too, Kernels.jl:99-145) when KHR_GRAPH=1 is set: kernels in the graph (0 = no graph in use) and replays so far. */
int32_t khr_graph_info(khr_ctx* ctx, int64_t* kernels_per_graph, int64_t* replays);
int32_t khr_set_profiling(khr_ctx* ctx, int32_t mode);
/* index in [0, count); pass out = NULL to query only the count */
int32_t khr_kernel_stat_get(khr_ctx* ctx, int32_t index, khr_kernel_stat* out, int32_t* count);
/* bytes of device memory held by the context */
int32_t khr_device_bytes(khr_ctx* ctx, int64_t* bytes);

#ifdef __cplusplus
}
#endif
#endif /* KHRONOS_B200_H */
