#!/usr/bin/env python
"""bench.py — Mcells/s per time step of the FDTD hot path (3-D Float32, PML + DFT).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload NAME]

A "step" is one full Khronos time step (H half-step, H-DFT, E half-step + ADE, E-DFT)
over the whole grid.  Default workload = BASELINE.json configs[1] (waveguide_mode,
480x240x132 Float32, PML + flux/mode DFT monitors); for N > 1 the same cell is
stacked N times along z (weak scaling, one z slab per GPU, halo over NCCL).
`value` = Nx*Ny*Nz*K / t / 1e6 with t the max over ranks of the CUDA-event time of
the K steps (reference definition, src/Simulation.jl:517-519).

--impl reference times the reference algorithm on the host cores: the reference is
Julia and cannot run in this image, so it is the C++/OpenMP port in oracle/
(cpu_baseline.kind == "port"), on a bounded sample of the same workload.
"""
import argparse
import json
import math
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np


# ----------------------------------------------------------------------------- helpers
def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            p = json.load(f)
        return float(p["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region by ONE long-running
    `nvidia-smi -lms` child (started before, stopped after): nothing forks inside the timed loops."""

    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.index, self.proc, self.rows = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return
        try:
            self.proc.terminate()
            out, _ = self.proc.communicate(timeout=5)
            self.rows = [[x.strip() for x in ln.split(",")] for ln in out.strip().splitlines() if ln.strip()]
        except Exception:
            try:
                self.proc.kill()
            except Exception:
                pass
        self.proc = None

    def summary(self):
        sm = [float(r[0]) for r in self.rows if len(r) >= 6 and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) >= 6 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(r) >= 6 and r[2 + i].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


def make_desc(name, nranks, sample_scale=1.0):
    from khronos_b200 import workloads as w
    if name == "waveguide_mode":
        res = max(4, int(round(40 * sample_scale)))
        return w.waveguide_mode(res=res, z_stack=nranks)
    if name == "sphere":
        return w.sphere(res=max(4, int(round(64 * sample_scale))))
    if name == "sphere256":
        return w.sphere(res=32)
    if name == "uled":
        return w.uled(res=max(4, int(round(40 * sample_scale))))
    if name.startswith("dipole"):
        n = int(name[6:] or 256)
        d = w.dipole(n)
        return d
    if name == "metalens":
        return w.metalens(nx=1024, ny=1024, nz=256 * nranks, res=32)
    if name == "metalens_full":
        # BASELINE.json configs[4]: 2048 x 2048 x 512 per GPU, ~20k rotated pillars (benchmark/metalens.jl);
        # only sensible with --rasterizer device (the numpy point sampler visits every voxel per object)
        return w.metalens(nx=2048, ny=2048, nz=512 * nranks, res=32, pillars=144, rotate=True)
    raise SystemExit("unknown workload " + name)


def bytes_per_cell_model(census, per_voxel_eps, w=4):
    """84 B (72 B scalar eps) + 80/112/144 B on 1/2/3-PML-axis voxels (Float32)."""
    tot = float(sum(census))
    base = (21 if per_voxel_eps else 18) * w
    return base + (census[1] * 20 * w + census[2] * 28 * w + census[3] * 36 * w) / tot


# ----------------------------------------------------------------------------- reference arm (CPU)
def cpu_run(name, steps, warmup, budget_s=40.0):
    """Time the oracle port on a bounded sample of the workload; returns (Mcells/s, info)."""
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
    # pick the sample: shrink the resolution until `steps` steps fit the budget (assume ~12 Mcells/s/core-ish)
    scale = 1.0
    est_rate = 8.0e6 * max(cores, 1)
    while True:
        d = make_desc(name, 1, scale)
        sim = w.build_simulation(d, np.float32)
        cells = sim.Nx * sim.Ny * sim.Nz
        if cells * (steps + warmup) / est_rate <= budget_s or scale <= 0.15:
            break
        scale *= 0.8
    o, _ = oracle_from_simulation(sim)
    o.step(max(warmup, 1))
    t0 = time.perf_counter()
    o.step(steps)
    dt = time.perf_counter() - t0
    rate = cells * steps / dt / 1e6
    info = {"value": rate, "unit": "Mcells/s", "cores": cores, "kind": "port",
            "sample": "%s at resolution scale %.2f: %dx%dx%d cells, %d steps, C++/OpenMP restatement of the "
                      "KernelAbstractions CPU path (oracle/)" % (d["name"], scale, sim.Nx, sim.Ny, sim.Nz, steps)}
    return rate, info, dt / steps * 1e3, d


def main_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    rate, info, ms, d = cpu_run(args.workload, args.steps, args.warmup, budget_s=90.0)
    line = {"impl": "reference", "metric": "Mcells/s per time step (3D Float32, PML+DFT)", "value": rate,
            "unit": "Mcells/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": args.workload, "sample": info["sample"]},
            "cpu_baseline": info,
            "e2e": {"value": rate, "unit": "Mcells/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


# ----------------------------------------------------------------------------- our arm (GPU)
def main_ours(args):
    import torch
    import khronos_b200 as kb
    from khronos_b200 import workloads as w

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
    n_gpus = world

    desc = make_desc(args.workload, n_gpus)
    dtype = np.float32 if args.dtype == "f32" else np.float64
    sim = w.build_simulation(desc, dtype, device=local_rank, rank=rank, nranks=n_gpus, rasterizer=args.rasterizer,
                             subpixel_smoothing=args.smoothing)
    t_prep = time.perf_counter()
    sim.prepare_simulation(comm_id=comm_id)
    t_prep = time.perf_counter() - t_prep
    cells = sim.Nx * sim.Ny * sim.Nz

    def barrier():
        sim.sync()
        if world > 1:
            torch.distributed.barrier()

    # ---- device-timed region: W warm-up + exactly K steps, inputs resident in HBM
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()   # samples through warm-up, the timed region and the e2e loop (all under load)
    sim.step(args.warmup)
    barrier()
    sim.step(args.steps)
    ms = sim.last_step_ms()           # CUDA events on the library's stream around the K steps (syncs)
    launches = sim.last_launches
    barrier()
    # second pass of the same K steps with one CUDA-event pair around every kernel launch and the
    # kernels of a half-step serialised on one stream (in the `value` pass they overlap on several
    # streams, where a per-kernel duration is not defined): per-kernel durations for the roofline
    sim.set_profiling(3)
    sim.step(args.steps)
    ms_profiled = sim.last_step_ms()
    barrier()
    stats = sim.kernel_stats()
    sim.set_profiling(0)
    if world > 1:
        t = torch.tensor([ms], device="cuda", dtype=torch.float64)
        torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
        ms = float(t.item())
    value = cells * args.steps / (ms * 1e-3) / 1e6

    # ---- end-to-end through the public API with host buffers in the timed region:
    # every step pushes the host-evaluated source amplitudes (h2d) and reads the DFT
    # convergence norms back (d2h), as run(sim; until_after_sources=stop_when_dft_decayed)
    # does in the reference (Simulation.jl:411-485); the monitors are read out at the end.
    e2e_steps = args.steps
    barrier()
    t0 = time.perf_counter()
    h2d = d2h = 0
    for _ in range(e2e_steps):
        sim.step(1)          # host evaluates a(t) per source, ships it in the kernel-parameter buffer
        h2d += 32 * len(sim.source_ids)
        sim.monitor_norms()  # stop_when_dft_decayed's per-step convergence metric (cached between DFT updates)
        d2h += 8 * len(sim.dft_monitors) / max(1, sim.dft_monitors[0].decimation if sim.dft_monitors else 1)
    # results the user reads after the run: get_flux of every flux monitor (reduced on the device,
    # nf doubles each; boxes split across ranks are read as arrays instead), Array(md.fields) of
    # every other DFT monitor (FluxMonitor.jl:92-102)
    out_bytes = 0
    for m in sim.monitors:
        if isinstance(m, kb.FluxMonitor) and world == 1:
            out_bytes += sim.get_flux(m).size * 8
        else:
            for dm in (m.monitors if isinstance(m, kb.FluxMonitor) else [m]):
                out_bytes += sim.get_dft(dm).size * 2 * np.dtype(dtype).itemsize
    barrier()
    e2e_s = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([e2e_s], device="cuda", dtype=torch.float64)
        torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
        e2e_s = float(t.item())
    e2e_value = cells * e2e_steps / e2e_s / 1e6
    if rank == 0:
        sampler.stop()
    if world > 1:
        torch.distributed.barrier()
        torch.distributed.destroy_process_group()

    if rank != 0:
        return
    peak, peak_src = peaks()
    census = sim.voxel_census()
    per_voxel_eps = sim.material_arrays["eps_inv"] is not None or args.rasterizer == "device"
    wbytes = np.dtype(dtype).itemsize
    bpc = bytes_per_cell_model(census, per_voxel_eps, wbytes)
    # dominant kernel = largest total CUDA-event time over the K steps of the serialised pass
    dom = max(stats, key=lambda s: s["total_ms"]) if stats else None
    roof = None
    kern_ms_total = sum(s["total_ms"] for s in stats)
    if dom and dom["launches"] > 0:
        avg_ms = dom["total_ms"] / dom["launches"]
        ach = dom["alg_bytes_per_launch"] / (avg_ms * 1e-3) / 1e9
        roof = {"bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak, "traffic": None,
                "kernel": dom["name"], "avg_launch_ms": avg_ms, "alg_bytes_per_launch": dom["alg_bytes_per_launch"],
                "share_of_step": dom["total_ms"] / kern_ms_total if kern_ms_total else None, "peak_source": peak_src,
                "whole_step": {"alg_bytes_per_cell": bpc, "achieved": value * 1e6 * bpc / n_gpus / 1e9,
                               "frac_of_measured": value * 1e6 * bpc / n_gpus / 1e9 / peak,
                               "frac_of_8TBs": value * 1e6 * bpc / n_gpus / 8.0e12}}
        tfile = os.path.join(ROOT, "profiles", "traffic_%s.json" % args.workload)
        if os.path.exists(tfile):
            try:
                tj = json.load(open(tfile))
                roof["traffic"] = tj.get(dom["name"])
                if roof["traffic"] is None and dom["name"].endswith(",muniform>"):
                    # capture taken before the constant-material PML tiles got their own table: same tiles
                    # (all but a handful), same loads skipped, launched then as part of the ",marr>" table
                    roof["traffic"] = tj.get(dom["name"].replace(",muniform>", ",marr>"))
                    roof["traffic_note"] = "ncu capture of the same tiles inside the former <...,marr> launch"
            except Exception:
                pass
        if roof["traffic"]:
            # the same launch against the DRAM bytes ncu counted for it (the kernels eliminate B/D and
            # skip constant-material loads, so they move less than the reference's algorithmic count)
            roof["traffic_achieved"] = roof["traffic"] / (avg_ms * 1e-3) / 1e9
            roof["traffic_frac"] = roof["traffic_achieved"] / peak
    cpu = None
    if n_gpus == 1 and not args.no_cpu:
        _, cpu, _, _ = cpu_run(args.workload, 6, 1, budget_s=15.0)
    line = {"metric": "Mcells/s per time step (3D Float32, PML+DFT)", "value": value, "unit": "Mcells/s",
            "n_gpus": n_gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": args.dtype, "data": "synthetic",
            "config": {"workload": desc["name"], "grid": [sim.Nx, sim.Ny, sim.Nz], "parallelism": "z-slab x%d" % n_gpus,
                       "pml_cells": int(round(desc["pml"][0][0] * desc["resolution"])),
                       "dft_monitors": len(sim.dft_monitors), "dft_decimation": sim.dft_monitors[0].decimation if sim.dft_monitors else None,
                       "voxel_census_0123_pml_axes": census, "device_bytes": sim.device_bytes(),
                       "l2": "working set %.0f MB > 126 MB L2, no flush needed" % (sim.device_bytes() / 1e6),
                       "prepare_s": t_prep, "rasterizer": args.rasterizer, "subpixel_smoothing": args.smoothing},
            "gpu_launches": int(launches),
            "e2e": {"value": e2e_value, "unit": "Mcells/s", "h2d_bytes_per_step": h2d / e2e_steps,
                    "d2h_bytes_per_step": (d2h + out_bytes) / e2e_steps,
                    "what": "sim.step(1) through the Python API/C ABI per step + host source amplitudes in + DFT norms out; get_flux / DFT arrays read at the end"},
            "roofline": roof, "cpu_baseline": cpu, "clocks": sampler.summary(),
            "ms_per_step_serialised_with_kernel_events": ms_profiled / args.steps,
            "kernels": [{k: s[k] for k in ("name", "launches", "total_ms", "ctas", "uniform_ctas")} for s in stats]}
    print(json.dumps(line))


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=1000)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours")
    ap.add_argument("--workload", default="waveguide_mode")
    ap.add_argument("--dtype", default="f32")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--rasterizer", default="host", choices=["host", "device"])
    ap.add_argument("--smoothing", default=None, choices=["volume", "anisotropic"])
    a = ap.parse_args()
    a.warmup = max(a.warmup, 3)
    if a.impl == "reference":
        main_reference(a)
    else:
        main_ours(a)
