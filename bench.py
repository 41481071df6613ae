#!/usr/bin/env python
"""bench.py — Mcells/s per time step of the FDTD hot path (3-D Float32, PML + DFT).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload NAME]

A "step" is one full Khronos time step (H half-step, H-DFT, E half-step + ADE, E-DFT) over the
whole grid.  Default workload = BASELINE.json configs[2], the largest single-GPU configuration and
the one SURVEY.md §8(d) works the roofline target on: dielectric sphere 512^3 Float32, 64-cell
PML, geometry rasterised on the device with the reference's anisotropic subpixel smoothing, 24
flux DFT planes x 21 frequencies.  For N > 1 the same cell is stacked N times along z (weak
scaling: one ball + flux box per GPU, one cost-balanced z slab per GPU, halo planes over NCCL).
`value` = Nx*Ny*Nz*K / t / 1e6 with t the max over ranks of the CUDA-event time of the K steps
(reference definition, src/Simulation.jl:517-519).  The timed window always contains at least one
DFT update step (the warm-up is extended by up to D-1 untimed steps so that a multiple of the
monitor decimation D falls inside it; `config.dft_updates_in_window`).

`extra` carries secondary lines measured in the same run: at N = 1 the other named configurations
(waveguide_mode 480x240x132, uled 280x280x100, dipole 500^3, metalens 2048x2048x512) and one Float64
line; at N > 1 the metalens slab decomposition (2048x2048x512 per GPU, BASELINE.json configs[4]) and a
`parity` record (a small multi-rank case against the CPU oracle).

--impl reference times the reference algorithm on the host cores: the reference is Julia and cannot
run in this image, so it is the C++/OpenMP port in oracle/ (cpu_baseline.kind == "port"), on a
bounded sample of the same workload.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np

METRIC = "Mcells/s per time step (3D Float32, PML+DFT)"


# ----------------------------------------------------------------------------- helpers
def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            p = json.load(f)
        return float(p["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """SM clock and throttle reasons sampled every ~5 ms through NVML by a background thread (ctypes calls
    into the library release the GIL), time-stamped so that the samples inside the timed window can be told
    from the rest.  Falls back to one long-running `nvidia-smi -lms 20` child when NVML cannot be loaded."""

    REASONS = ((0x8, "hw_slowdown"), (0x40, "hw_thermal_slowdown"), (0x20, "sw_thermal_slowdown"), (0x4, "sw_power_cap"))

    def __init__(self, index=0):
        self.index, self.rows, self.thread, self.stop_flag, self.proc, self.t_smi0 = index, [], None, False, None, 0.0
        self.max_mhz = None

    def _loop(self, nv, h):
        while not self.stop_flag:
            try:
                mhz = nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)
                try:
                    rs = nv.nvmlDeviceGetCurrentClocksEventReasons(h)
                except Exception:
                    rs = nv.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                self.rows.append((time.perf_counter(), float(mhz), int(rs)))
            except Exception:
                pass
            time.sleep(0.004)

    def start(self):
        try:
            import pynvml as nv
            nv.nvmlInit()
            h = nv.nvmlDeviceGetHandleByIndex(self.index)
            self.max_mhz = float(nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM))
            self.thread = threading.Thread(target=self._loop, args=(nv, h), daemon=True)
            self.thread.start()
        except Exception:
            self.thread = None
            try:
                q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
                     "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
                self.t_smi0 = time.perf_counter()
                self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q, "--format=csv,noheader,nounits",
                                              "-lms", "20"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            except Exception:
                self.proc = None

    def stop(self):
        self.stop_flag = True
        if self.thread is not None:
            self.thread.join(timeout=2)
        if self.proc is not None:
            try:
                t1 = time.perf_counter()
                self.proc.terminate()
                out, _ = self.proc.communicate(timeout=5)
                lines = [[x.strip() for x in ln.split(",")] for ln in out.strip().splitlines() if ln.strip()]
                n = max(len(lines), 1)
                for i, r in enumerate(lines):
                    if len(r) >= 6 and r[0].replace(".", "").isdigit():
                        rs = sum(bit for (bit, _), v in zip(self.REASONS, r[2:6]) if v.lower().startswith("active"))
                        self.max_mhz = max(self.max_mhz or 0.0, float(r[1]))
                        self.rows.append((self.t_smi0 + (t1 - self.t_smi0) * (i + 0.5) / n, float(r[0]), rs))
            except Exception:
                pass
            self.proc = None

    def summary(self, t0, t1):
        inside = [r for r in self.rows if t0 <= r[0] <= t1]
        use = inside if inside else self.rows
        bits = 0
        for r in use:
            bits |= r[2]
        return {"sm_mhz": float(np.median([r[1] for r in use])) if use else None, "sm_max_mhz": self.max_mhz,
                "reasons": [n for b, n in self.REASONS if bits & b], "samples": len(inside), "samples_total": len(self.rows),
                "how": "NVML every ~5 ms from a background thread; `samples` are those inside the timed window" if self.thread
                       else "nvidia-smi -lms 20 child"}


WORKLOAD_DEFAULTS = {
    # name: (rasterizer, subpixel smoothing)
    "sphere": ("device", "anisotropic"), "sphere256": ("device", "anisotropic"), "metalens_full": ("device", None),
    "metalens": ("device", None), "periodic_bloch": ("device", None), "metalens_strong": ("device", None),
}
# how a workload grows with the number of ranks: "weak" = the cell is stacked N x along z, "strong" = fixed total size
WORKLOAD_SCALING = {"sphere": "weak", "sphere256": "weak", "waveguide_mode": "weak", "metalens": "weak", "metalens_full": "weak"}


def make_desc(name, nranks, sample_scale=1.0):
    from khronos_b200 import workloads as w
    if name == "waveguide_mode":
        res = max(4, int(round(40 * sample_scale)))
        return w.waveguide_mode(res=res, z_stack=nranks)
    if name == "sphere":
        return w.sphere(res=max(4, int(round(64 * sample_scale))), z_stack=nranks)
    if name == "sphere256":
        return w.sphere(res=32, z_stack=nranks)
    if name == "uled":
        return w.uled(res=max(4, int(round(40 * sample_scale))))
    if name.startswith("dipole"):
        n = int(name[6:] or 256)
        return w.dipole(max(16, int(round(n * sample_scale / 8)) * 8) if sample_scale != 1.0 else n)
    if name == "periodic_bloch":
        # benchmark/periodic_bloch.jl at a size that loads a B200: 16 x 16 holes at resolution 40 (640 x 640 x 180, complex fields)
        n = 16 if sample_scale == 1.0 else 3
        return w.periodic_bloch(res=max(8, int(round(40 * sample_scale))), n_cells=n)
    if name == "metalens":
        return w.metalens(nx=1024, ny=1024, nz=256 * nranks, res=32)
    if name == "metalens_full":
        # BASELINE.json configs[4]: 2048 x 2048 x 512 per GPU, ~20k rotated pillars (benchmark/metalens.jl);
        # rasterised on the device (the numpy point sampler visits every voxel per object)
        if sample_scale != 1.0:
            n = max(64, int(round(2048 * sample_scale / 32)) * 32)
            return w.metalens(nx=n, ny=n, nz=max(64, n // 4), res=32, pillars=max(2, int(144 * sample_scale)), rotate=True)
        return w.metalens(nx=2048, ny=2048, nz=512 * nranks, res=32, pillars=144, rotate=True)
    if name == "metalens_strong":
        # the fixed 2048 x 2048 x 512 metalens split over the ranks (strong scaling)
        return w.metalens(nx=2048, ny=2048, nz=512, res=32, pillars=144, rotate=True)
    raise SystemExit("unknown workload " + name)


def config_of(desc, sim, n_gpus, rasterizer, smoothing, dtype):
    """The workload as both arms name it (host-side facts only, identical for --impl ours / reference)."""
    wb = np.dtype(dtype).itemsize
    fields_mb = 6.0 * (sim.Nx + 2) * (sim.Ny + 2) * (sim.Nz + 2) * wb / 1e6
    return {"workload": desc["name"], "grid": [sim.Nx, sim.Ny, sim.Nz], "parallelism": "z-slab x%d" % n_gpus,
            "pml_cells": int(round(max(p[0] for p in desc["pml"]) * desc["resolution"])), "dft_monitors": len(sim.dft_monitors),
            "rasterizer": rasterizer, "subpixel_smoothing": smoothing,
            "l2": "inputs larger than L2: the six field arrays alone are %.0f MB per step >> 126 MB, no flush needed" % fields_mb}


# ----------------------------------------------------------------------------- reference arm (CPU)
def cpu_run(name, steps, warmup, budget_s=40.0):
    """Time the oracle port on a bounded sample of the workload; returns (Mcells/s, info, ms/step)."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import oracle as ko
    from bridge import oracle_from_simulation
    from khronos_b200 import workloads as w
    ko.build()
    # all the host threads this process may use (torchrun exports OMP_NUM_THREADS=1 to its workers)
    try:
        avail = len(os.sched_getaffinity(0))
    except Exception:
        avail = os.cpu_count() or 1
    ko.set_num_threads(avail)
    cores = ko.num_threads()
    rast, smooth = WORKLOAD_DEFAULTS.get(name, ("host", None))
    # pick the sample: shrink the resolution until `steps` steps fit the budget (assume ~8 Mcells/s per core)
    scale = 1.0
    est_rate = 8.0e6 * max(cores, 1)
    while True:
        d = make_desc(name, 1, scale)
        sim = w.build_simulation(d, np.float32, rasterizer=rast, subpixel_smoothing=smooth)
        cells = sim.Nx * sim.Ny * sim.Nz
        if (cells * (steps + warmup) / est_rate <= budget_s and cells <= 70e6) or scale <= 0.15:
            break
        scale *= 0.8
    o, _ = oracle_from_simulation(sim, check=False)
    o.step(max(warmup, 1))
    t0 = time.perf_counter()
    o.step(steps)
    dt = time.perf_counter() - t0
    rate = cells * steps / dt / 1e6
    info = {"value": rate, "unit": "Mcells/s", "cores": cores, "kind": "port",
            "sample": "%s at resolution scale %.2f: %dx%dx%d cells, %d steps, C++/OpenMP restatement of the "
                      "KernelAbstractions CPU path (oracle/)" % (d["name"], scale, sim.Nx, sim.Ny, sim.Nz, steps)}
    return rate, info, dt / steps * 1e3


def main_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from khronos_b200 import workloads as w
    rast, smooth = args.rasterizer, args.smoothing
    desc = make_desc(args.workload, args.gpus)
    full = w.build_simulation(desc, np.float32, rasterizer=rast, subpixel_smoothing=smooth)
    full_dms = sum(len(m.monitors) if hasattr(m, "monitors") else 1 for m in full.monitors)
    full.dft_monitors = [None] * full_dms      # host-side count only; nothing is prepared at full size
    cfg = config_of(desc, full, args.gpus, rast, smooth, np.float32)
    rate, info, ms = cpu_run(args.workload, args.steps, args.warmup, budget_s=90.0)
    line = {"impl": "reference", "metric": METRIC, "value": rate, "unit": "Mcells/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic", "config": cfg, "cpu_baseline": info,
            "e2e": {"value": rate, "unit": "Mcells/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


# ----------------------------------------------------------------------------- our arm (GPU)
def aligned_warmup(warmup, steps, decs):
    """Smallest W' >= warmup such that a DFT update step (timestep % D == 0) falls inside [W', W'+steps)."""
    decs = sorted(set(int(d) for d in decs if d and d > 1))
    if not decs:
        return warmup
    D = decs[-1]
    if steps >= D:
        return warmup
    w = warmup
    while (w + steps // 2) % D != 0:
        w += 1
    return w


def run_workload(name, args, ctx, steps, warmup, dtype=np.float32, e2e=True, sampler=None):
    """One workload on the ranks of this job.  Returns the result dict (rank 0) or None."""
    import torch
    import khronos_b200 as kb
    from khronos_b200 import workloads as w
    world, rank, local_rank, comm_id = ctx["world"], ctx["rank"], ctx["local_rank"], ctx["comm_id"]
    rast, smooth = WORKLOAD_DEFAULTS.get(name, ("host", None))
    if name == args.workload:
        rast, smooth = args.rasterizer, args.smoothing
    desc = make_desc(name, world)
    sim = w.build_simulation(desc, dtype, device=local_rank, rank=rank, nranks=world, rasterizer=rast, subpixel_smoothing=smooth,
                             slab_rule=args.slab_rule)
    t_prep = time.perf_counter()
    sim.prepare_simulation(comm_id=comm_id)
    t_prep = time.perf_counter() - t_prep
    cells = sim.Nx * sim.Ny * sim.Nz

    def barrier():
        sim.sync()
        if world > 1:
            torch.distributed.barrier()

    def allmax(x):
        if world == 1:
            return x
        t = torch.tensor([x], device="cuda", dtype=torch.float64)
        torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
        return float(t.item())

    def gather(x):
        if world == 1:
            return [x]
        t = torch.tensor([x], device="cuda", dtype=torch.float64)
        out = [torch.zeros_like(t) for _ in range(world)]
        torch.distributed.all_gather(out, t)
        return [float(v.item()) for v in out]

    # ---- device-timed region: W' warm-up + exactly K steps, inputs resident in HBM
    decs = [m.decimation for m in sim.dft_monitors]
    w_al = aligned_warmup(warmup, steps, decs)
    sim.step(w_al)
    barrier()
    t_win0 = time.perf_counter()
    sim.step(steps)
    ms_rank = sim.last_step_ms()      # CUDA events on the library's stream around the K steps (syncs)
    t_win1 = time.perf_counter()
    launches = sim.last_launches
    barrier()
    ms = allmax(ms_rank)
    ms_ranks = gather(ms_rank)
    value = cells * steps / (ms * 1e-3) / 1e6
    dft_updates = sum(1 for t in range(w_al, w_al + steps) for D in set(decs) if t % max(D, 1) == 0)
    # ---- per-kernel durations: the same K steps with the kernels of a half-step serialised on one stream and one
    # CUDA-event pair around every launch (in the `value` pass they overlap on several streams, where a per-kernel
    # duration is not defined)
    sim.set_profiling(3)
    sim.step(steps)
    ms_profiled = sim.last_step_ms()
    barrier()
    stats = sim.kernel_stats()
    halo = None
    if world > 1:
        # halo wait measured with the normal multi-stream overlap (events around the wait only matter here)
        sim.set_profiling(2)
        sim.step(steps)
        barrier()
        wms, nex = sim.comm_stats()
        halo = gather(wms / steps)
    sim.set_profiling(0)

    # ---- end-to-end through the public API with host buffers in the timed region: every step is one call
    # through Python / ctypes / the C ABI, the host evaluates the source amplitudes and ships them with the
    # launch (h2d), and reads the DFT convergence norms back (d2h) as run(sim; until_after_sources =
    # stop_when_dft_decayed) does in the reference (Simulation.jl:411-485); get_flux of every flux monitor
    # (reduced on the device, across ranks with one all-reduce) and the other DFT arrays are read at the end.
    e2e_rec = None
    if e2e:
        # first-call work stays outside the window like any warm-up: the scratch of the flux reduction and, on
        # several ranks, NCCL's lazy set-up of its all-reduce channels
        for m in sim.monitors[:1]:
            if isinstance(m, kb.FluxMonitor):
                sim.get_flux(m)
        sim.monitor_norms()
        barrier()
        t0 = time.perf_counter()
        h2d = d2h = 0
        for _ in range(steps):
            sim.step(1)
            h2d += 32 * len(sim.source_ids)
            sim.monitor_norms()
            d2h += 8 * len(sim.dft_monitors) / max(1, decs[0] if decs else 1)
        out_bytes = 0
        for m in sim.monitors:
            if isinstance(m, kb.FluxMonitor):
                out_bytes += sim.get_flux(m).size * 8
            else:
                out_bytes += sim.get_dft(m).size * 2 * np.dtype(dtype).itemsize
        barrier()
        e2e_s = allmax(time.perf_counter() - t0)
        e2e_rec = {"value": cells * steps / e2e_s / 1e6, "unit": "Mcells/s", "h2d_bytes_per_step": h2d / steps,
                   "d2h_bytes_per_step": (d2h + out_bytes) / steps,
                   "what": "sim.step(1) per step through the Python API / C ABI: host source amplitudes in, DFT norms out; "
                           "get_flux (device reduction) / DFT arrays read at the end"}
    census = sim.voxel_census()
    dev_bytes = sim.device_bytes()
    gk, gr = sim.graph_info()
    cfg = config_of(desc, sim, world, rast, smooth, dtype)
    sim.close()
    if rank != 0:
        return None
    peak, peak_src = peaks()
    dom = max(stats, key=lambda s: s["total_ms"]) if stats else None
    kern_ms_total = sum(s["total_ms"] for s in stats)
    roof = None
    if dom and dom["launches"] > 0:
        avg_ms = dom["total_ms"] / dom["launches"]
        ach = dom["alg_bytes_per_launch"] / (avg_ms * 1e-3) / 1e9
        step_bytes = sum(s["alg_bytes_per_launch"] for s in stats)        # every table is launched once per step
        ref_bytes = sum(s["ref_model_bytes_per_launch"] for s in stats)
        step_gbs = step_bytes / (ms_rank / steps * 1e-3) / 1e9
        roof = {"bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak, "traffic": None,
                "kernel": dom["name"], "avg_launch_ms": avg_ms, "alg_bytes_per_launch": dom["alg_bytes_per_launch"],
                "bytes_model": "compulsory bytes of this implementation per launch (9w/12w per voxel, constant-material tiles 9w, "
                               "+4w per PML axis, + conductivity / ADE arrays where present), DESIGN.md §4",
                "bytes_vs_reference_model": dom["alg_bytes_per_launch"] / dom["ref_model_bytes_per_launch"]
                if dom["ref_model_bytes_per_launch"] else None,
                "share_of_step": dom["total_ms"] / kern_ms_total if kern_ms_total else None, "peak_source": peak_src,
                "ctas": dom["ctas"],
                "whole_step": {"alg_bytes_per_cell": step_bytes / (cells / world), "reference_model_bytes_per_cell": ref_bytes / (cells / world),
                               "achieved": step_gbs, "frac": step_gbs / peak,
                               "note": "sum of the compulsory bytes of every launch of a step / device time of the step on this rank"}}
        tfile = os.path.join(ROOT, "profiles", "traffic_%s.json" % name)
        if world == 1 and os.path.exists(tfile):
            # DRAM bytes ncu counted for this kernel (one --set full capture, per launch, scripts/ncu_summary.py).  Used only
            # when the capture is of the same launch: same kernel AND same launch shape.  The persistent TMA kernel always
            # launches one CTA per SM, so its shape is identified by the number of work items of the table instead.
            try:
                tj = json.load(open(tfile))
                ent = next((v for k, v in tj.items() if dom["name"].startswith(k)), None)
                persistent = dom["name"].startswith("halfstep_tma_kernel")
                same = isinstance(ent, dict) and (int(ent.get("items", -1)) == int(dom["ctas"]) if persistent
                                                  else int(ent.get("grid", -1)) == int(dom["ctas"]))
                if same:
                    roof["traffic"] = ent["bytes"]
                    roof["traffic_source"] = "profiles/traffic_%s.json (ncu --set full, dram__bytes_read.sum + dram__bytes_write.sum per launch)" % name
                    roof["traffic_achieved"] = ent["bytes"] / (avg_ms * 1e-3) / 1e9
                    roof["traffic_frac"] = roof["traffic_achieved"] / peak
            except Exception:
                pass
    res = {"value": value, "unit": "Mcells/s", "ms_per_step": ms / steps, "steps": steps, "warmup": w_al, "dtype": "f32" if dtype is np.float32 else "f64",
           "config": cfg, "gpu_launches": int(launches), "e2e": e2e_rec, "roofline": roof,
           "details": {"voxel_census_0123_pml_axes": census, "device_bytes": dev_bytes, "prepare_s": t_prep,
                       "dft_decimation": decs[0] if decs else None, "dft_updates_in_window": dft_updates,
                       "warmup_requested": warmup, "warmup_run": w_al, "slabs": [list(s) for s in sim.slabs],
                       "cuda_graph": {"kernels_per_step_graph": gk, "replays": gr,
                                      "host_launches_per_step": 1 if gk else None},
                       "ms_per_step_serialised_with_kernel_events": ms_profiled / steps,
                       "kernels": [{k: s[k] for k in ("name", "launches", "total_ms", "ctas", "uniform_ctas", "alg_bytes_per_launch")} for s in stats]},
           "window": (t_win0, t_win1)}
    if world > 1:
        res["per_rank"] = {"ms_per_step": [m / steps for m in ms_ranks], "halo_wait_ms_per_step": halo}
    return res


def main_ours(args):
    import torch
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the product has no CPU path (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    comm_id = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        from khronos_b200 import distributed as kd
        comm_id = kd.broadcast_unique_id(rank)
    ctx = dict(world=world, rank=rank, local_rank=local_rank, comm_id=comm_id)
    dtype = np.float32 if args.dtype == "f32" else np.float64

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    main = run_workload(args.workload, args, ctx, args.steps, args.warmup, dtype)
    clocks = None
    if rank == 0:
        clocks = sampler.summary(*main["window"])

    extra = {}
    name_of = {}
    if not args.no_extra:
        def fresh_id():
            # every workload gets its own communicator (the context owns it)
            if world == 1:
                return None
            from khronos_b200 import distributed as kd
            return kd.broadcast_unique_id(rank)

        def short(r):
            if r is None:
                return None
            rf = r["roofline"] or {}
            return {"value": r["value"], "unit": "Mcells/s", "ms_per_step": r["ms_per_step"], "steps": r["steps"], "dtype": r["dtype"],
                    "config": r["config"], "e2e": (r["e2e"] or {}).get("value"), "gpu_launches": r["gpu_launches"],
                    "roofline": {k: rf.get(k) for k in ("kernel", "frac", "achieved", "share_of_step", "bytes_vs_reference_model", "whole_step")},
                    "dft_updates_in_window": r["details"]["dft_updates_in_window"], "per_rank": r.get("per_rank"),
                    "scaling": WORKLOAD_SCALING.get(name_of[id(r)], "strong") if world > 1 else None,
                    "slabs": r["details"]["slabs"] if world > 1 else None}
        if world == 1:
            plan = [("waveguide_mode", np.float32, 400), ("uled", np.float32, 400), ("dipole500", np.float32, 60),
                    ("metalens_full", np.float32, 20), ("sphere", np.float64, 70)]
        else:
            # the other named shapes on N ranks: stacked cells (weak) or the fixed domain cut into N slabs (strong)
            plan = [("metalens_full", np.float32, 20), ("metalens_strong", np.float32, 40), ("waveguide_mode", np.float32, 200),
                    ("dipole500", np.float32, 60), ("uled", np.float32, 200)]
        for name, dt_, k in plan:
            key = name + ("_f64" if dt_ is np.float64 else "")
            if name == args.workload and dt_ is dtype:
                continue
            try:
                ctx["comm_id"] = fresh_id()
                r = run_workload(name, args, ctx, k, 5, dt_)
                if r is not None:
                    name_of[id(r)] = name
                extra[key] = short(r)
            except Exception as e:          # an extra line must never take the headline down
                extra[key] = {"error": str(e)[:300]}
        if world > 1:
            sys.path.insert(0, os.path.join(ROOT, "scripts"))
            try:
                from mgpu_parity import run_case
                # the same kernels the headline uses: the persistent TMA half-step kernel next to the halo exchange
                extra["parity"] = run_case(rank, world, local_rank, fresh_id(), flags=("--tma",))
            except Exception as e:
                extra["parity"] = {"error": str(e)[:300], "ok": False}
    if rank == 0:
        sampler.stop()
    if world > 1:
        torch.distributed.barrier()
        torch.distributed.destroy_process_group()
    if rank != 0:
        return
    cpu = None
    if world == 1 and not args.no_cpu:
        _, cpu, _ = cpu_run(args.workload, 6, 1, budget_s=15.0)
    main.pop("window")
    line = {"metric": METRIC, "value": main["value"], "unit": "Mcells/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": main["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": args.dtype, "data": "synthetic", "config": main["config"], "gpu_launches": main["gpu_launches"],
            "e2e": main["e2e"], "roofline": main["roofline"], "cpu_baseline": cpu, "clocks": clocks,
            "details": main["details"], "extra": extra}
    if world > 1:
        line["per_rank"] = main["per_rank"]
    print(json.dumps(line))


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours")
    ap.add_argument("--workload", default="sphere")
    ap.add_argument("--dtype", default="f32")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-extra", action="store_true")
    ap.add_argument("--rasterizer", default=None, choices=["host", "device"])
    ap.add_argument("--smoothing", default=None, choices=["none", "volume", "anisotropic"])
    ap.add_argument("--slab-rule", default="cost", choices=["cost", "reference"])
    a = ap.parse_args()
    a.warmup = max(a.warmup, 3)
    d_r, d_s = WORKLOAD_DEFAULTS.get(a.workload, ("host", None))
    a.rasterizer = a.rasterizer or d_r
    a.smoothing = d_s if a.smoothing is None else (None if a.smoothing == "none" else a.smoothing)
    if a.rasterizer != "device":
        a.smoothing = None
    if a.impl == "reference":
        main_reference(a)
    else:
        main_ours(a)
