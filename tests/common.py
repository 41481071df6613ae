"""Shared builders: the same configuration handed to the CPU oracle and to the product."""
import numpy as np

import khronos_b200 as kb
from bridge import oracle_from_simulation  # oracle/bridge.py (test infrastructure)


def rel_l2(a, b):
    a = np.asarray(a)
    b = np.asarray(b)
    a = a.astype(np.complex128 if np.iscomplexobj(a) else np.float64)
    b = b.astype(np.complex128 if np.iscomplexobj(b) else np.float64)
    den = np.sqrt(np.sum(np.abs(b) ** 2))
    num = np.sqrt(np.sum(np.abs(a - b) ** 2))
    return num / den if den > 0 else num


class Pair:
    """One configuration instantiated twice: oracle (CPU) and product (GPU)."""

    def __init__(self, cell, res, pml, dtype=np.float32, courant=0.5, sources=(), monitors=(), eps_inv=None,
                 mu_inv=None, sigma_D=None, sigma_B=None, poles=(), absorbers=None, center=(0.0, 0.0, 0.0),
                 build_gpu=True, rank=0, nranks=1, comm_id=None, device=0, geometry=None,
                 boundary_conditions=None, chi3=None, grid_spacing=None, rasterizer="host", subpixel_smoothing=None):
        self.dtype = dtype
        bnd = None if pml is None else [[p, p] if np.isscalar(p) else list(p) for p in pml]
        ksrc = [kb.UniformSource(tp, comp, c, s) for (comp, c, s, tp) in sources]
        kmon = [kb.DFTMonitor(comp, c, s, f, dec) for (comp, c, s, f, dec) in monitors]
        self.kmon = kmon
        self.k = kb.Simulation(cell, list(center), res, ksrc, boundaries=bnd, monitors=kmon, Courant=courant,
                               dtype=dtype, eps_inv=eps_inv, mu_inv=mu_inv, sigma_D=sigma_D, sigma_B=sigma_B,
                               poles=list(poles), absorbers=absorbers, rank=rank, nranks=nranks, device=device,
                               geometry=geometry, boundary_conditions=boundary_conditions, chi3=chi3, grid_spacing=grid_spacing,
                               rasterizer=rasterizer, subpixel_smoothing=subpixel_smoothing)
        self.grid = self.k.grid
        # the oracle always simulates the whole domain (single chunk)
        whole = self.k if nranks == 1 else kb.Simulation(
            cell, list(center), res, ksrc, boundaries=bnd, monitors=kmon, Courant=courant, dtype=dtype, eps_inv=eps_inv,
            mu_inv=mu_inv, sigma_D=sigma_D, sigma_B=sigma_B, poles=list(poles), absorbers=absorbers, geometry=geometry,
            boundary_conditions=boundary_conditions, chi3=chi3, grid_spacing=grid_spacing,
            rasterizer=rasterizer, subpixel_smoothing=subpixel_smoothing)
        self.o, self.omon = oracle_from_simulation(whole)
        self.build_gpu = build_gpu
        if build_gpu:
            self.k.prepare_simulation(comm_id=comm_id)

    def step(self, n):
        self.o.step(n)
        if self.build_gpu:
            self.k.step(n)
            self.k.sync()

    def field_errors(self):
        return {comp: rel_l2(self.k.get_field(comp), self.o.get_field(comp)) for comp in range(6)}

    def total_field_error(self):
        num = den = 0.0
        for comp in range(6):
            a = self.k.get_field(comp).astype(np.float64)
            b = self.o.get_field(comp).astype(np.float64)
            num += np.sum((a - b) ** 2)
            den += np.sum(b ** 2)
        return np.sqrt(num / den)
