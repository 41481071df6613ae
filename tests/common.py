"""Shared builders: the same configuration handed to the CPU oracle and to the product."""
import numpy as np

import khronos_b200 as kb
import oracle as ko  # oracle/oracle.py (test infrastructure)

DT = {np.float32: 0, np.float64: 1}


def rel_l2(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    den = np.sqrt(np.sum(np.abs(b) ** 2))
    num = np.sqrt(np.sum(np.abs(a - b) ** 2))
    return num / den if den > 0 else num


class Pair:
    """One configuration instantiated twice: oracle (CPU) and product (GPU)."""

    def __init__(self, cell, res, pml, dtype=np.float32, courant=0.5, sources=(), monitors=(), eps_inv=None,
                 mu_inv=None, sigma_D=None, sigma_B=None, poles=(), absorbers=None, center=(0.0, 0.0, 0.0),
                 build_gpu=True, rank=0, nranks=1, comm_id=None, device=0):
        self.dtype = dtype
        bnd = None if pml is None else [[p, p] if np.isscalar(p) else list(p) for p in pml]
        self.o = ko.OracleSim(dtype, list(cell), list(center), res, courant, bnd)
        g = kb.Grid(cell, center, res, courant, dtype)
        self.grid = g
        ksrc = []
        for (comp, c, s, tp) in sources:
            ksrc.append(kb.UniformSource(tp, comp, c, s))
        kmon = [kb.DFTMonitor(comp, c, s, f, dec) for (comp, c, s, f, dec) in monitors]
        self.kmon = kmon
        if build_gpu:
            self.k = kb.Simulation(cell, list(center), res, ksrc, boundaries=bnd, monitors=kmon, Courant=courant,
                                   dtype=dtype, eps_inv=eps_inv, mu_inv=mu_inv, sigma_D=sigma_D, sigma_B=sigma_B,
                                   poles=list(poles), absorbers=absorbers, rank=rank, nranks=nranks, device=device)
        else:
            self.k = None
        # ---- oracle side: same index maps computed by the product's host code
        helper = kb.Simulation(cell, list(center), res, ksrc, boundaries=bnd, dtype=dtype, absorbers=absorbers)
        arrays = {"eps_inv": eps_inv, "mu_inv": mu_inv, "sigma_D": sigma_D, "sigma_B": sigma_B}
        arrays = {k: (None if v is None else [np.array(x, dtype=dtype) for x in v]) for k, v in arrays.items()}
        arrays = helper._apply_absorbers(arrays)
        pl = [(w, gm, helper._zero_pole_sigma_in_pml(np.asarray(s, dtype=dtype))) for (w, gm, s) in poles]
        if pl:
            import math
            chi1 = np.zeros(tuple(g.N), dtype=dtype)
            for (w0, gam, s) in pl:
                dtd = float(g.dt)
                g1i = 1.0 / (1.0 + gam * math.pi * dtd)
                c = dtype(g1i * (gam * 2 * math.pi * dtd * dtd) / 2) if w0 == 0.0 else dtype(g1i * ((2 * math.pi * w0 * dtd) ** 2) / 2)
                chi1 = chi1 + s * c
            if arrays["eps_inv"] is None:
                arrays["eps_inv"] = [np.full(tuple(g.N), dtype(1), dtype=dtype) for _ in range(3)]
            nz = chi1 != 0
            for d in range(3):
                e = arrays["eps_inv"][d].copy()
                e[nz] = e[nz] / (dtype(1) + e[nz] * chi1[nz])
                arrays["eps_inv"][d] = e
        for key in ("eps_inv", "mu_inv", "sigma_D", "sigma_B"):
            if arrays[key] is not None:
                for d in range(3):
                    self.o.set_material_array(key, d, arrays[key][d])
        for (w0, gam, s) in pl:
            self.o.add_pole(w0, gam, s)
        for (comp, c, s, tp) in sources:
            start, end = g.grid_volume(c, s, comp)
            dims = [end[a] - start[a] + 1 for a in range(3)]
            amp = helper._source_amplitude(kb.UniformSource(tp, comp, c, s), comp, start, dims)
            ct = np.complex64 if dtype is np.float32 else np.complex128
            self.o.add_source(comp, start, amp.astype(ct), tp.kind, tp.params(dtype))
        self.omon = []
        dec = helper_decimation(sources, g)
        for (comp, c, s, f, d) in monitors:
            start, end = g.grid_volume(c, s, comp)
            dd = d if d != 1 else dec
            self.omon.append(self.o.add_dft(comp, start, end, [float(dtype(x)) for x in f], dd))
        self.o.prepare("single")
        if self.k is not None:
            self.k.prepare_simulation(comm_id=comm_id)

    def step(self, n):
        self.o.step(n)
        if self.k is not None:
            self.k.step(n)
            self.k.sync()

    def field_errors(self):
        out = {}
        for comp in range(6):
            out[comp] = rel_l2(self.k.get_field(comp), self.o.get_field(comp))
        return out

    def total_field_error(self):
        num = den = 0.0
        for comp in range(6):
            a = self.k.get_field(comp).astype(np.float64)
            b = self.o.get_field(comp).astype(np.float64)
            num += np.sum((a - b) ** 2)
            den += np.sum(b ** 2)
        return np.sqrt(num / den)


def helper_decimation(sources, g):
    import math
    f_max = 0.0
    for (_, _, _, tp) in sources:
        f_max = max(f_max, tp.f_max())
    if f_max <= 0:
        return 1
    return max(1, int(math.floor(1.0 / (2.0 * f_max * float(g.dt)))))
