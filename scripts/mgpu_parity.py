"""Multi-GPU parity: N ranks (torchrun, one per GPU, z slabs + NCCL halo exchange) against the
single-domain CPU oracle on rank 0.

  torchrun --nproc-per-node N scripts/mgpu_parity.py [--periodic] [--kerr] [--nonuniform] [--reference-slabs] [--bloch] [--blochz] [--tma] [--thin]

`run_case` is also what `bench.py --gpus N` calls for its `parity` sub-record and what
tests/test_gpu_multirank.py spawns under torch.distributed.run.
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")]
import numpy as np


def run_case(rank, world, local_rank, comm_id, flags=(), nsteps=120):
    """A 44x40x96 Float32 case that exercises the rank seams: random per-voxel eps, a Drude slab
    straddling the middle of the z axis, an E point source and an H sheet source, a DFT box spanning all
    ranks, a flux plane (x-normal, spans all ranks) reduced on the device across ranks.  Every rank
    steps its slab; fields are gathered and DFT boxes summed on the host; rank 0 compares with the
    single-domain oracle.  Collective over the default torch.distributed group.  Returns a dict on rank 0."""
    import torch.distributed as dist
    import khronos_b200 as kb
    from khronos_b200 import distributed as kd

    rng = np.random.default_rng(1234)
    # --thin: 36 z cells, 10 of them PML at either end -> with 4 ranks every slab is 9 planes and the end slabs lie
    # entirely inside the z PML (what the 280x280x100 uled cell looks like on 8 ranks)
    thin = "--thin" in flags
    nz = 36 if thin else 96
    N = (44, 40, nz)
    eps = [(1.0 / rng.uniform(1.0, 4.0, N)).astype(np.float32) for _ in range(3)]
    sg = np.zeros(N, dtype=np.float32)
    sg[:, :, nz // 2 - 8:nz // 2 + 8] = 1.5          # a Drude slab that straddles the rank boundary for 2 ranks
    srcs = [kb.UniformSource(kb.ContinuousWaveSource(1.0), kb.EZ, [0, 0, 0.05], [0, 0, 0]),
            kb.UniformSource(kb.ContinuousWaveSource(1.2), kb.HY, [0.3, 0, -0.5 if thin else -1.0], [1.0, 1.0, 0])]
    Lz = nz / 10.0
    fm = kb.FluxMonitor([0.6, 0, 0], [0, 2.0, 3.0 if thin else 7.0], [1.0, 1.1], 2)     # x-normal plane through every slab
    mons = [kb.DFTMonitor(kb.EX, [0, 0, 0], [0, 3, Lz], [1.0, 1.2], 2),
            kb.DFTMonitor(kb.HZ, [0, 0.2, 0.3 if thin else 1.0], [3, 0, 2.4 if thin else 6.0], [1.0], 1), fm]
    kw = dict(boundaries=[[1.0, 1.0]] * 3, monitors=mons, eps_inv=eps, poles=[(0.0, 0.3, sg)])
    if "--periodic" in flags:
        # x and z periodic (z closes the halo ring across ranks), PML on y only
        kw["boundaries"] = [[0.0, 0.0], [1.0, 1.0], [0.0, 0.0]]
        kw["boundary_conditions"] = [[kb.Periodic(), kb.Periodic()], [kb.PML(), kb.PML()], [kb.Periodic(), kb.Periodic()]]
    bloch = "--bloch" in flags or "--blochz" in flags
    if bloch:
        # Bloch boundaries (complex fields, two real contexts per rank): x with k = 1.3 always; --blochz adds a
        # Bloch-periodic z axis, i.e. the halo ring across the ranks with the phase on the seam ghosts
        zb = "--blochz" in flags
        kw["boundaries"] = [[0.0, 0.0], [1.0, 1.0], [0.0, 0.0] if zb else [1.0, 1.0]]
        kw["boundary_conditions"] = [[kb.Bloch(1.3), kb.Bloch(1.3)], [kb.PML(), kb.PML()],
                                     [kb.Bloch(-0.7), kb.Bloch(-0.7)] if zb else [kb.PML(), kb.PML()]]
    if "--kerr" in flags:
        # a Kerr block that straddles the rank boundary and overlaps the Drude slab partly
        chi3 = np.zeros(N, dtype=np.float32)
        chi3[14:30, 12:28, 30:50] = 0.3   # chi3 |E|^2 up to ~0.04: 2.0 gives 0.25, where Float32 round-off is amplified past 1e-5
        kw["chi3"] = chi3
    if "--nonuniform" in flags:
        # graded spacing along x and z (z is the decomposed axis: every rank gets its slice)
        ix, iz = np.arange(N[0]), np.arange(N[2])
        kw["grid_spacing"] = [(0.1 * (1 + 0.25 * np.sin(2 * np.pi * ix / N[0]))).astype(np.float32), None,
                              (0.1 * (1 + 0.2 * np.cos(2 * np.pi * iz / N[2] + 0.7))).astype(np.float32)]
    if "--tma" in flags:
        os.environ["KHR_TMA"] = "1"      # the persistent TMA half-step kernel next to the halo exchange (automatic only on large slabs)
    rule = "reference" if "--reference-slabs" in flags else "cost"
    sim = kb.Simulation([4.4, 4.0, Lz], [0, 0, 0], 10, srcs, rank=rank, nranks=world, device=local_rank, slab_rule=rule, **kw)
    sim.prepare_simulation(comm_id=comm_id)
    sim.step(nsteps)
    sim.sync()
    fields = [kd.gather_fields(sim, c) for c in range(6)]
    fields_im = [kd.gather_fields(sim, c, part="imag") for c in range(6)] if bloch else None
    dfts = [kd.reduce_dft(sim, m) for m in sim.dft_monitors[:2]]
    flux = sim.get_flux(fm)                    # collective: ncclAllReduce of the four accumulators, then the device reduction
    slabs = sim.slabs
    sim.close()
    out = None
    if rank == 0:
        from bridge import oracle_from_simulation
        whole = kb.Simulation([4.4, 4.0, Lz], [0, 0, 0], 10, srcs, **kw)
        o, mids = oracle_from_simulation(whole)
        o.step(nsteps)
        num = den = 0.0
        for c in range(6):
            b = o.get_field(c)
            num += ((fields[c].astype(np.float64) - b) ** 2).sum()
            den += (b ** 2).sum()
            if bloch:
                bi = o.get_field(c, "imag")
                num += ((fields_im[c].astype(np.float64) - bi) ** 2).sum()
                den += (bi ** 2).sum()
        err = float((num / den) ** 0.5)
        derr = [float(np.linalg.norm(a - o.get_dft(m)) / np.linalg.norm(o.get_dft(m))) for a, m in zip(dfts, mids[:2])]
        fref = o.flux(fm.normal, mids[2:6])
        ferr = float(np.linalg.norm(flux - fref) / np.linalg.norm(fref))
        out = dict(case="44x40xNZ f32, eps per voxel, Drude slab across the seam, 2 sources, 2 DFT boxes + 1 flux plane, %d steps%s"
                        .replace("NZ", str(nz)) % (nsteps, (" " + " ".join(flags)) if flags else ""),
                   world=world, slabs=[list(s) for s in slabs], field_rel_l2=err, dft_rel_l2=derr, flux_rel_l2=ferr,
                   tolerance=1e-5, ok=bool(err < 1e-5 and max(derr) < 1e-5 and ferr < 1e-5))
    return out


def main():
    import torch
    import torch.distributed as dist
    from khronos_b200 import distributed as kd
    rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(lr)
    dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
    cid = kd.broadcast_unique_id(rank)
    flags = [a for a in sys.argv[1:] if a.startswith("--")]
    res = run_case(rank, world, lr, cid, flags)
    ok = True
    if rank == 0:
        print("mgpu parity %s" % res)
        ok = res["ok"]
        sys.stdout.flush()
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.broadcast(flag, 0)
    dist.barrier()
    dist.destroy_process_group()
    if not flag.item():
        sys.exit(1)


if __name__ == "__main__":
    main()
