"""Process-group plumbing for the z-slab decomposition (one process per GPU).

Mirrors src/Distributed.jl:25-72: the reference bootstraps NCCL by creating the
unique id on rank 0 and broadcasting it with MPI; here torch.distributed carries
the 128 bytes (NCCL or gloo backend), the halo traffic itself is ncclSend/ncclRecv
issued by libkhronos_b200.so on its own communicator.
"""
import ctypes as C

from . import _lib


def unique_id():
    buf = (C.c_char * 128)()
    _lib.check(_lib.lib().khr_comm_unique_id(buf))
    return bytes(buf)


def broadcast_unique_id(rank, src=0):
    """Rank `src` creates the ncclUniqueId; every rank returns the same 128 bytes."""
    import torch
    import torch.distributed as dist

    dev = "cuda" if dist.get_backend() == "nccl" else "cpu"
    if rank == src:
        t = torch.tensor(list(unique_id()), dtype=torch.uint8, device=dev)
    else:
        t = torch.zeros(128, dtype=torch.uint8, device=dev)
    dist.broadcast(t, src=src)
    return bytes(t.cpu().tolist())


def gather_fields(sim, comp, part="real"):
    """Visualization.jl:294-333 _pull_fields_from_device: global field from the per-rank slabs
    (part = "real" | "imag" | "complex" for complex fields)."""
    import numpy as np
    import torch.distributed as dist

    local = sim.get_field(comp, part) if part != "real" else sim.get_field(comp)
    if sim.nranks == 1:
        return local
    parts = [None] * sim.nranks
    dist.all_gather_object(parts, local)
    return np.concatenate(parts, axis=2)


def reduce_dft(sim, monitor):
    """Sum of the per-rank (zero-padded) DFT boxes, like the reference's MPI.Reduce! (Visualization.jl:328)."""
    import torch.distributed as dist

    local = sim.get_dft(monitor)
    if sim.nranks == 1:
        return local
    parts = [None] * sim.nranks
    dist.all_gather_object(parts, local)
    return sum(parts)
