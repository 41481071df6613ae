"""Synthetic restatements of the reference's benchmark configurations.

Each builder returns a plain description (dict) that the host mirror turns into
a `Simulation`; bench.py and the tests build the CPU oracle from the same dict.
Geometry is point sampled (GeometryPrimitives.jl / the mode solver are un-vendored
and out of scope), fields start from zero and evolve under the source.

  dipole          benchmark/dipole.jl:33-51
  waveguide_mode  benchmark/waveguide_mode.jl:41-90   (BASELINE.json configs[1])
  sphere          benchmark/sphere.jl:34-72
  uled            benchmark/uled.jl:43-215
  metalens        benchmark/metalens.jl:44-140
"""
import math

import numpy as np

from .grid import EX, EY, EZ, HX, HY, HZ
from .simulation import (PML, Ball, Bloch, ContinuousWaveSource, Cuboid, Cylinder, DFTMonitor, DrudeSusceptibility, FluxMonitor,
                         LorentzianSusceptibility, Material, Object, Simulation, UniformSource)


def dipole(n=40, res=None, pml=1.0, monitors=True):
    """n^3 cells. The reference rows are size_xyz*res = 40..500 with PML 1.0 (= res cells)."""
    res = res or {40: 10, 80: 20, 160: 20, 320: 40, 500: 50, 256: 32, 512: 64}.get(n, 10)
    L = n / res
    mons = [DFTMonitor(EZ, [0, 0, 0], [L, L, 0], [0.9, 1.0, 1.1])] if monitors else []
    return dict(name="dipole_%d" % n, cell_size=[L, L, L], resolution=res, pml=[[pml, pml]] * 3, courant=0.5,
                sources=[UniformSource(ContinuousWaveSource(1.0), EZ, [0, 0, 0], [0, 0, 0])], monitors=mons,
                geometry=[])


def _gauss_profile(cy, cz, wy, wz):
    def f(pt, comp):
        return np.exp(-(((pt[1] - cy) / wy) ** 2) - (((pt[2] - cz) / wz) ** 2)) + 0 * pt[0]
    return f


def waveguide_mode(res=40, wg_length=10.0, z_stack=1):
    """480x240x132 at res 40. `z_stack` repeats the cell along z (weak-scaling runs)."""
    eps_si, eps_sio2 = 3.47 ** 2, 1.44 ** 2
    pml = 1.0
    cell = [wg_length + 2 * pml, 4.0 + 2 * pml, (6 * 0.22 + 2 * pml) * z_stack]
    geom = [Object(Cuboid([0, 0, 0], [wg_length + 4.0, 0.5, 0.22]), Material(epsilon=eps_si)),
            Object(Cuboid([0, 0, 0], [1e9, 1e9, 1e9]), Material(epsilon=eps_sio2))]
    f0 = 1.0 / 1.55
    cw = ContinuousWaveSource(f0)
    xs = -wg_length / 2 + 1.0
    prof = _gauss_profile(0.0, 0.0, 0.35, 0.2)
    # mode source = 4 tangential components on the plane (mode profile -> separable Gaussian)
    srcs = [UniformSource(cw, c, [xs, 0, 0], [0, 2.0, 2.0], amplitude=a, profile=prof)
            for c, a in ((EY, 1.0), (EZ, 0.1), (HY, -0.1), (HZ, 1.0))]
    freqs = list(np.linspace(f0 * 0.9, f0 * 1.1, 11))
    mons = [FluxMonitor([xs + 1.0, 0, 0], [0, 2.0, 2.0], freqs), FluxMonitor([wg_length / 2 - 1.0, 0, 0], [0, 2.0, 2.0], freqs),
            FluxMonitor([0.0, 0, 0], [0, 2.0, 2.0], [f0])]
    return dict(name="waveguide_mode_res%d" % res, cell_size=cell, resolution=res, pml=[[pml, pml]] * 3, courant=0.5,
                sources=srcs, monitors=mons, geometry=geom)


def sphere(res=64, radius=2.5, nfreq=21, z_stack=1):
    """512^3 at res 64: eps=3 ball, Ex+Hy plane sources, 6-face flux box x 21 frequencies.
    `z_stack` repeats the cell (ball + flux box) along z: weak-scaling runs, one cell per GPU."""
    s = 2.0 + 1.0 + 2 * radius
    z0 = -s * z_stack / 2 + 1.0
    inf = float("inf")
    cw = ContinuousWaveSource(1.0)
    srcs = [UniformSource(cw, EX, [0, 0, z0], [inf, inf, 0.0]), UniformSource(cw, HY, [0, 0, z0], [inf, inf, 0.0])]
    b = radius + 0.25
    freqs = list(np.linspace(0.8, 1.2, nfreq))
    mons, geom = [], []
    for q in range(z_stack):
        zc = (-z_stack / 2 + 0.5 + q) * s
        geom.append(Object(Ball([0, 0, zc], radius), Material(epsilon=3.0)))
        for axis in range(3):
            for sgn in (-1, 1):
                c = [0.0, 0.0, zc]
                c[axis] += sgn * b
                sz = [2 * b] * 3
                sz[axis] = 0.0
                mons.append(FluxMonitor(c, sz, freqs))
    name = "sphere_res%d" % res + ("" if z_stack == 1 else "_x%d" % z_stack)
    return dict(name=name, cell_size=[s, s, s * z_stack], resolution=res, pml=[[1.0, 1.0]] * 3, courant=0.5,
                sources=srcs, monitors=mons, geometry=geom)


def uled(res=40, lorentz=True):
    """280x280x100 at res 40: layered stack with an Ag Drude layer, dipole, near-to-far DFT box."""
    cell = [7.0, 7.0, 2.5]
    pml = 0.5
    ag = [DrudeSusceptibility(gamma=0.05, sigma=60.0)]
    if lorentz:
        ag.append(LorentzianSusceptibility(omega_0=3.0, gamma=0.4, sigma=1.2))
    geom = [Object(Cuboid([0, 0, -0.55], [1e9, 1e9, 0.1]), Material(epsilon=1.0, susceptibilities=ag)),
            Object(Cuboid([0, 0, -0.3], [1e9, 1e9, 0.4]), Material(epsilon=6.0)),
            Object(Cuboid([0, 0, 0.05], [1e9, 1e9, 0.3]), Material(epsilon=5.3)),
            Object(Cuboid([0, 0, -0.9], [1e9, 1e9, 0.6]), Material(epsilon=2.1))]
    f0 = 1.0 / 0.45
    srcs = [UniformSource(ContinuousWaveSource(f0), EY, [0, 0, -0.25], [0, 0, 0])]
    freqs = list(np.linspace(f0 * 0.95, f0 * 1.05, 5))
    mons = []
    bx, bz = 2.6, 0.6
    faces = [([0, 0, bz], [2 * bx, 2 * bx, 0]), ([-bx, 0, 0.2], [0, 2 * bx, 0.8]), ([bx, 0, 0.2], [0, 2 * bx, 0.8]),
             ([0, -bx, 0.2], [2 * bx, 0, 0.8]), ([0, bx, 0.2], [2 * bx, 0, 0.8])]
    for c, sz in faces:
        mons.append(FluxMonitor(c, sz, freqs))
    return dict(name="uled_res%d" % res, cell_size=cell, resolution=res, pml=[[pml, pml]] * 3, courant=0.5,
                sources=srcs, monitors=mons, geometry=geom)


def metalens(nx=512, ny=512, nz=128, res=32, pml_cells=15, pillars=8, rotate=False):
    """Slab-decomposed metalens-like domain: substrate half-space + pillar array, plane source,
    focal-plane DFT. The benchmark size is 2048x2048x512 per GPU; smaller sizes are parity cases."""
    cell = [nx / res, ny / res, nz / res]
    pml = pml_cells / res
    zsub = -cell[2] / 2 + 0.35 * cell[2]
    geom = []
    pitch = cell[0] / (pillars + 1)
    rng = np.random.default_rng(1234)
    for i in range(pillars):
        for j in range(pillars):
            w = pitch * rng.uniform(0.3, 0.7)
            c = [-cell[0] / 2 + pitch * (i + 1), -cell[1] / 2 + cell[1] / (pillars + 1) * (j + 1), zsub + 0.3]
            axes = None
            if rotate:  # benchmark/metalens.jl rotates every pillar about z
                th = rng.uniform(0, math.pi)
                axes = [[math.cos(th), math.sin(th), 0.0], [-math.sin(th), math.cos(th), 0.0], [0.0, 0.0, 1.0]]
            geom.append(Object(Cuboid(c, [w, w, 0.6], axes=axes), Material(epsilon=5.76)))
    geom.append(Object(Cuboid([0, 0, zsub - cell[2]], [1e9, 1e9, 2 * cell[2]]), Material(epsilon=2.13)))
    cw = ContinuousWaveSource(1.0 / 0.64)
    inf = float("inf")
    zs = -cell[2] / 2 + pml + 4.0 / res
    srcs = [UniformSource(cw, c, [0, 0, zs], [inf, inf, 0.0], amplitude=a)
            for c, a in ((EX, 1.0), (EY, 0.5), (HX, -0.5), (HY, 1.0))]
    zf = cell[2] / 2 - pml - 6.0 / res
    mons = [DFTMonitor(EX, [0, 0, zf], [cell[0], cell[1], 0], [1.0 / 0.64]),
            DFTMonitor(EY, [0, 0, zf], [cell[0], cell[1], 0], [1.0 / 0.64])]
    return dict(name="metalens_%dx%dx%d" % (nx, ny, nz), cell_size=cell, resolution=res, pml=[[pml, pml]] * 3,
                courant=0.55, sources=srcs, monitors=mons, geometry=geom)


def periodic_bloch(res=20, n_cells=3):
    """benchmark/periodic_bloch.jl:41-110: n x n supercell of air holes (Cylinder) in an eps = 12 slab, Bloch(k) in
    x (X point) and y (k = 0), PML in z, Hz point dipole; complex fields.  Reference rows: res 20 / n 1 (20x20x90),
    res 20 / n 3 (60x60x90), res 30 / n 3."""
    a, r_hole, t_slab = 1.0, 0.2, 0.5
    pml_z, buffer_z = 1.0, 1.0
    cell_z = t_slab + 2 * buffer_z + 2 * pml_z
    cell_xy = n_cells * a
    geom = []
    for ix in range(n_cells):
        for iy in range(n_cells):
            geom.append(Object(Cylinder([(ix + 0.5) * a - cell_xy / 2, (iy + 0.5) * a - cell_xy / 2, 0.0], r_hole, t_slab, [0, 0, 1]),
                               Material(epsilon=1.0)))
    geom.append(Object(Cuboid([0, 0, 0], [cell_xy + 1.0, cell_xy + 1.0, t_slab]), Material(epsilon=12.0)))
    kx = 0.5 * 2 * math.pi / a
    srcs = [UniformSource(ContinuousWaveSource(0.3), HZ, [0.13 * a - cell_xy / 2, 0.27 * a - cell_xy / 2, 0.0], [0, 0, 0])]
    mons = [DFTMonitor(HZ, [0, 0, 0], [cell_xy, cell_xy, 0], [0.3])]
    return dict(name="periodic_bloch_res%d_n%d" % (res, n_cells), cell_size=[cell_xy, cell_xy, cell_z], resolution=res,
                pml=[[0.0, 0.0], [0.0, 0.0], [pml_z, pml_z]], courant=0.5, sources=srcs, monitors=mons, geometry=geom,
                boundary_conditions=[[Bloch(kx), Bloch(kx)], [Bloch(0.0), Bloch(0.0)], [PML(), PML()]])


WORKLOADS = {"dipole": dipole, "waveguide_mode": waveguide_mode, "sphere": sphere, "uled": uled, "metalens": metalens,
             "periodic_bloch": periodic_bloch}


def build_simulation(desc, dtype=np.float32, device=0, rank=0, nranks=1, rasterizer="host", subpixel_smoothing=None,
                     slab_rule="cost"):
    return Simulation(desc["cell_size"], [0.0, 0.0, 0.0], desc["resolution"], desc["sources"], boundaries=desc["pml"],
                      geometry=desc["geometry"], monitors=desc["monitors"], Courant=desc["courant"], dtype=dtype,
                      device=device, rank=rank, nranks=nranks, rasterizer=rasterizer,
                      subpixel_smoothing=subpixel_smoothing, slab_rule=slab_rule,
                      boundary_conditions=desc.get("boundary_conditions"))
