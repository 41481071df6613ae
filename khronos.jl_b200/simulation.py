"""Host-side mirror of the Khronos `Simulation` / `run(sim; until=...)` / monitor API.

The reference host is Julia (no toolchain in this image), so the part of
Khronos that sits *in front of* the time-step hot path is mirrored here in
Python with the reference's names and argument meaning: `Simulation(...)`,
`prepare_simulation`, `step`, `run(until=... | until_after_sources=...)`,
`run_benchmark`, `UniformSource`, `ContinuousWaveSource`, `GaussianPulseSource`,
`DFTMonitor`, `FluxMonitor`, `get_flux`, `Material`, `Object`, susceptibilities,
`Absorber`.  Everything per-step is delegated to libkhronos_b200.so through the
C ABI (`_lib.py`); nothing here computes fields on the host.

Reference: src/DataStructures.jl:714-801, src/Simulation.jl:92-382, 492-527.
"""
import ctypes as C
import math

import numpy as np

from . import _lib
from . import chunking
from .grid import CENTER, EX, EY, EZ, HX, HY, HZ, Grid, component_stagger, interpolation_weight

# --------------------------------------------------------------------------
# time profiles (src/Sources/TimeSources.jl)
# --------------------------------------------------------------------------


class ContinuousWaveSource:
    """TimeSources.jl:48-68. a(t) = exp(-i 2π fcen t); never cuts off."""

    def __init__(self, fcen):
        self.fcen = float(fcen)

    kind = _lib.TIME_CW

    def params(self, T):
        return [float(T(self.fcen)), 0.0, 0.0, 0.0]

    def f_max(self, T=np.float64):
        return float(T(self.fcen))     # the profile stores fcen as backend_number (TimeSources.jl:52-58)

    def cutoff(self):
        return None


class GaussianPulseSource:
    """TimeSources.jl:74-136 (constructor arithmetic reproduced in Float64)."""

    kind = _lib.TIME_GAUSSIAN

    def __init__(self, fcen, fwidth, start_time=0.0, cutoff_scale=5.0):
        self.ctor_args = (float(fcen), float(fwidth), float(start_time), float(cutoff_scale))
        width = 1.0 / fwidth
        cutoff = width * cutoff_scale + start_time
        self.fwidth = math.sqrt(-2.0 * math.log(1e-7)) / (width * math.pi)
        while math.exp(-cutoff * cutoff / (2 * width * width)) < 1e-100:
            cutoff *= 0.9
        period = 1.0 / fcen
        # Julia round(): ties to even, like Python's round()
        self.peak_time = round((cutoff / 2) / period) * period
        self.fcen, self.width, self.cutoff_ = float(fcen), width, cutoff

    def params(self, T):
        return [float(T(self.fcen)), float(T(self.width)), float(T(self.peak_time)), float(T(self.cutoff_))]

    def f_max(self, T=np.float64):
        return float(T(self.fcen)) + float(T(self.fwidth)) / 2   # Monitors.jl:43-45 on the T-typed fields

    def cutoff(self):
        return self.cutoff_


class CustomSource:
    """TimeSources.jl:139-183: amplitude supplied by a host callable every step."""

    kind = _lib.TIME_HOST

    def __init__(self, src_func, fcen, fwidth, end_time):
        self.src_func, self.fcen, self.fwidth, self.end_time = src_func, float(fcen), float(fwidth), float(end_time)

    def params(self, T):
        return [float(T(self.fcen)), float(T(self.fwidth)), 0.0, float(T(self.end_time))]

    def f_max(self, T=np.float64):
        return float(T(self.fcen)) + float(T(self.fwidth)) / 2

    def cutoff(self):
        return self.end_time


# --------------------------------------------------------------------------
# spatial sources (src/Sources/SpatialSources.jl:99-155, Sources.jl:40-135)
# --------------------------------------------------------------------------


class UniformSource:
    def __init__(self, time_profile, component, center, size, amplitude=1.0, profile=None):
        self.time_profile = time_profile
        self.components = [component] if np.isscalar(component) else list(component)
        self.center = [float(v) for v in center]
        self.size = [float(v) for v in size]
        self.amplitude = complex(amplitude)
        # optional spatial profile f(point, component) (stands in for the mode /
        # plane-wave / Gaussian-beam profiles, whose construction is out of scope)
        self.profile = profile


# --------------------------------------------------------------------------
# materials / geometry (inputs of the hot path; rasterised by point sampling)
# --------------------------------------------------------------------------


class LorentzianSusceptibility:
    """Susceptibility.jl:24-28."""

    def __init__(self, omega_0, gamma, sigma):
        self.omega_0, self.gamma, self.sigma = float(omega_0), float(gamma), float(sigma)


def DrudeSusceptibility(gamma, sigma):
    """Susceptibility.jl:39-40."""
    return LorentzianSusceptibility(0.0, gamma, sigma)


class Material:
    def __init__(self, epsilon=1.0, mu=1.0, sigma_D=0.0, sigma_B=0.0, susceptibilities=None, chi3=None):
        self.epsilon, self.mu = float(epsilon), float(mu)
        self.sigma_D, self.sigma_B = float(sigma_D), float(sigma_B)
        self.susceptibilities = list(susceptibilities or [])
        self.chi3 = None if chi3 is None else float(chi3)  # Kerr coefficient (DataStructures.jl:271)


class Ball:
    def __init__(self, center, radius):
        self.center, self.radius = [float(v) for v in center], float(radius)

    def contains(self, X, Y, Z):
        c = self.center
        return (X - c[0]) ** 2 + (Y - c[1]) ** 2 + (Z - c[2]) ** 2 <= self.radius ** 2


class Cuboid:
    """GeometryPrimitives Cuboid(c, d, axes): `axes` rows are the (orthogonal) edge directions."""

    def __init__(self, center, size, axes=None):
        self.center, self.size = [float(v) for v in center], [float(v) for v in size]
        self.axes = None
        if axes is not None:
            a = np.asarray(axes, dtype=np.float64).reshape(3, 3)
            self.axes = a / np.linalg.norm(a, axis=1, keepdims=True)

    def contains(self, X, Y, Z):
        c, s = self.center, self.size
        if self.axes is None:
            return (np.abs(X - c[0]) <= s[0] / 2) & (np.abs(Y - c[1]) <= s[1] / 2) & (np.abs(Z - c[2]) <= s[2] / 2)
        d = [X - c[0], Y - c[1], Z - c[2]]
        ok = True
        for k in range(3):
            p = (self.axes[k, 0] * d[0] + self.axes[k, 1] * d[1]) + self.axes[k, 2] * d[2]
            ok = ok & (np.abs(p) <= s[k] / 2)
        return ok


class Cylinder:
    """GeometryPrimitives Cylinder(c, r, h, a): radius r, height h along the axis a (normalised here)."""

    def __init__(self, center, radius, height, axis=(0.0, 0.0, 1.0)):
        self.center, self.radius, self.height = [float(v) for v in center], float(radius), float(height)
        a = np.asarray(axis, dtype=np.float64)
        self.axis = a / np.linalg.norm(a)

    def contains(self, X, Y, Z):
        c, a = self.center, self.axis
        d = [X - c[0], Y - c[1], Z - c[2]]
        p = (d[0] * a[0] + d[1] * a[1]) + d[2] * a[2]
        q = [d[0] - p * a[0], d[1] - p * a[1], d[2] - p * a[2]]
        return (np.abs(p) <= self.height / 2) & (((q[0] * q[0] + q[1] * q[1]) + q[2] * q[2]) <= self.radius ** 2)


class Object:
    def __init__(self, shape, material):
        self.shape, self.material = shape, material


# boundary conditions (DataStructures.jl:150-161)
class PML:
    code = 0


class Periodic:
    code = 1


class Bloch:
    """Bloch(k) (DataStructures.jl:158-160): wrap-around with the phase exp(i k L); any Bloch side
    switches the simulation to complex fields (Fields.jl:140-159)."""
    code = 1

    def __init__(self, k=0.0):
        self.k = float(k)


class PECBoundary:
    code = 2


class PMCBoundary:
    code = 3


class Absorber:
    """Boundaries.jl:17-21."""

    def __init__(self, num_layers=40, sigma_order=3, sigma_max=0.0):
        self.num_layers, self.sigma_order, self.sigma_max = int(num_layers), int(sigma_order), float(sigma_max)


# --------------------------------------------------------------------------
# monitors (src/Monitors/Monitors.jl:202-272, FluxMonitor.jl:18-70)
# --------------------------------------------------------------------------


class DFTMonitor:
    def __init__(self, component, center, size, frequencies, decimation=1):
        self.component = component
        self.center = [float(v) for v in center]
        self.size = [float(v) for v in size]
        self.frequencies = [float(f) for f in frequencies]
        self.user_decimation = int(decimation)   # as given; `decimation` may be raised by auto_decimate!
        self.decimation = int(decimation)
        self.id = None
        self.start = self.end = None


class FluxMonitor:
    """Four tangential DFT monitors on a plane (FluxMonitor.jl:18-70)."""

    def __init__(self, center, size, frequencies, decimation=1):
        self.center = [float(v) for v in center]
        self.size = [float(v) for v in size]
        self.frequencies = [float(f) for f in frequencies]
        self.decimation = int(decimation)
        zero = [i for i, s in enumerate(self.size) if s == 0.0]
        if len(zero) != 1:
            raise ValueError("FluxMonitor needs exactly one zero-size axis (the normal)")
        self.normal = zero[0]
        t1 = 1 if self.normal == 0 else 0
        t2 = 1 if self.normal == 2 else 2
        self.tangential = (t1, t2)
        self.monitors = [
            DFTMonitor(EX + t1, center, size, frequencies, decimation),
            DFTMonitor(EX + t2, center, size, frequencies, decimation),
            DFTMonitor(HX + t1, center, size, frequencies, decimation),
            DFTMonitor(HX + t2, center, size, frequencies, decimation),
        ]


class Near2FarMonitor(FluxMonitor):
    """Near2FarMonitor (DataStructures.jl:441-455, init_near2far_monitor Near2Far.jl:1137-1200): the
    same four tangential DFT monitors as a flux plane plus the far-field description."""

    def __init__(self, center, size, frequencies, observation_points=None, theta=None, phi=None, r=1e6,
                 normal_dir="+", medium_eps=1.0, medium_mu=1.0, decimation=1):
        super().__init__(center, size, frequencies, decimation)
        self.observation_points = None if observation_points is None else np.asarray(observation_points, dtype=np.float64)
        self.theta = None if theta is None else np.asarray(theta, dtype=np.float64)
        self.phi = None if phi is None else np.asarray(phi, dtype=np.float64)
        self.r = float(r)
        self.normal_sign = 1.0 if normal_dir in ("+", 1, +1.0) else -1.0
        self.medium_eps, self.medium_mu = float(medium_eps), float(medium_mu)


class DiffractionMonitor(FluxMonitor):
    """DiffractionMonitor (init_diffraction_monitor, DiffractionMonitor.jl:18-80): four tangential DFT
    monitors on a plane, decomposed into diffraction orders afterwards."""


class ModeMonitor(FluxMonitor):
    """ModeMonitor (DataStructures.jl:512-519): four tangential DFT monitors on a plane; the mode
    profiles themselves come from the caller (the mode solver is outside the hot path)."""


# --------------------------------------------------------------------------
# Simulation
# --------------------------------------------------------------------------


class Simulation:
    """Khronos.Simulation(; cell_size, cell_center, resolution, sources, boundaries, ...).

    Extra keywords that have no Julia counterpart: `dtype` (the reference selects
    precision with choose_backend, load_deps.jl:23-64), `device`, and the slab
    placement `rank` / `nranks` (the reference reads them from MPI,
    Distributed.jl:25-55).  Per-voxel material arrays may be handed in directly
    (`eps_inv`, `mu_inv`, `sigma_D`, `sigma_B`: dense (Nx,Ny,Nz) arrays per
    component, i.e. what init_geometry would have produced).
    """

    def __init__(self, cell_size, cell_center, resolution, sources, boundaries=None, absorbers=None, geometry=None,
                 monitors=None, Courant=0.5, dtype=np.float32, device=0, rank=0, nranks=1, eps_inv=None, mu_inv=None,
                 sigma_D=None, sigma_B=None, poles=None, boundary_conditions=None, chi3=None, grid_spacing=None,
                 rasterizer="host", subpixel_smoothing=None, slab_rule="cost"):
        # slab_rule: how the z axis is cut into one slab per rank — "cost" (equal work, default) or
        # "reference" (the literal index rule of the distributed PML-grid planner), chunking.z_slab_partition
        self.slab_rule = slab_rule
        # grid_spacing: [Δx, Δy, Δz], each None (uniform) or one spacing per cell — the reference's
        # Simulation(Δx = vector, ...) (DataStructures.jl:737-739)
        self.grid = Grid(cell_size, cell_center, resolution, Courant, dtype, spacing=grid_spacing)
        # boundary_conditions (DataStructures.jl:725): per axis [minus, plus] of PML / Periodic /
        # Bloch / PECBoundary / PMCBoundary instances (or classes); None == PML everywhere
        self.bc_codes = None
        self.periodic = [False, False, False]
        self.complex_fields = False          # _needs_complex_fields (Fields.jl:140-159)
        self.bloch_k = [0.0, 0.0, 0.0]
        if boundary_conditions is not None:
            codes = []
            for a, axis_bcs in enumerate(boundary_conditions):
                for bc in axis_bcs:
                    if isinstance(bc, Bloch):
                        self.complex_fields = True
                    codes.append(int(bc.code))
                # Chunking.jl:1739-1744: k of the first side if it is Bloch, else of the second
                if isinstance(axis_bcs[0], Bloch):
                    self.bloch_k[a] = axis_bcs[0].k
                elif len(axis_bcs) > 1 and isinstance(axis_bcs[1], Bloch):
                    self.bloch_k[a] = axis_bcs[1].k
            if len(codes) != 6:
                raise ValueError("boundary_conditions must be three [minus, plus] pairs")
            self.bc_codes = codes
            # wrap-around needs both sides periodic (Chunking.jl:1730-1733)
            self.periodic = [codes[2 * a] == 1 and codes[2 * a + 1] == 1 for a in range(3)]
        self.T = self.grid.T
        self.sources = list(sources or [])
        self.boundaries = None if boundaries is None else [[float(a), float(b)] for a, b in boundaries]
        self.absorbers = absorbers
        self.geometry = list(geometry or [])
        self.monitors = list(monitors or [])
        self.device, self.rank, self.nranks = int(device), int(rank), int(nranks)
        self.user_arrays = {"eps_inv": eps_inv, "mu_inv": mu_inv, "sigma_D": sigma_D, "sigma_B": sigma_B}
        self.user_poles = list(poles or [])  # (omega_0, gamma, sigma_array) triples
        self.user_chi3 = chi3                # dense (Nx,Ny,Nz) Kerr coefficient on the centre grid
        # rasterizer="device": eps^-1 / mu^-1 / sigma_D / sigma_B of `geometry` are produced on the GPU
        # (khr_geometry_rasterize), with the reference's subpixel smoothing if asked for
        # (subpixel_smoothing = None | "volume" | "anisotropic"; DataStructures.jl:66-76, :727).
        # rasterizer="host" is the point sampler below (no smoothing), which also serves the oracle.
        if rasterizer not in ("host", "device"):
            raise ValueError("rasterizer must be 'host' or 'device'")
        if subpixel_smoothing not in (None, "volume", "anisotropic"):
            raise ValueError("subpixel_smoothing must be None, 'volume' or 'anisotropic'")
        if subpixel_smoothing is not None and rasterizer != "device":
            raise ValueError("subpixel smoothing is implemented by the device rasterizer (rasterizer='device')")
        self.rasterizer, self.subpixel_smoothing = rasterizer, subpixel_smoothing
        self.ctx = None
        self.is_prepared = False
        self.dft_monitors = []
        self.timestep = 0
        self.Nx, self.Ny, self.Nz = self.grid.N
        self.dx, self.dy, self.dz = self.grid.dl
        self.dt = self.grid.dt

    # -------------------------------------------------------------- helpers
    def _coords(self, comp):
        """Coordinates of cells 1..N of a component grid (Geometry.jl _precompute_coords)."""
        o = self.grid.component_origin(comp)
        out = []
        for a in range(3):
            if self.grid.dlv[a] is None:
                # origin + (i - 1) * Δ with the Int * T product formed in T (Geometry.jl:351-375)
                out.append(o[a] + (np.arange(self.grid.N[a]).astype(self.T) * self.T(self.grid.dl[a])).astype(np.float64))
            else:  # _build_coords(Δ::AbstractVector, ...) (Geometry.jl:377-395): origin + sum(Δ[1:i-1])
                cum = np.concatenate([[0.0], np.cumsum(self.grid.dlv[a].astype(np.float64))[:-1]])
                out.append(o[a] + cum)
        return out

    def _rasterize(self):
        """Point-sampled stand-in for init_geometry (Geometry.jl:450-663): per Yee component,
        cell i takes the material of the first object containing the component position."""
        T = self.T
        g = self.grid
        arrays = {k: None for k in ("eps_inv", "mu_inv", "sigma_D", "sigma_B", "chi3")}
        poles = []
        if not self.geometry:
            return arrays, poles
        host = self.rasterizer == "host"     # the device rasterizer produces these four kinds itself
        need_eps = host and any(o.material.epsilon != 1.0 for o in self.geometry)
        need_mu = host and any(o.material.mu != 1.0 for o in self.geometry)
        need_sd = host and any(o.material.sigma_D != 0.0 for o in self.geometry)
        need_sb = host and any(o.material.sigma_B != 0.0 for o in self.geometry)
        shape = tuple(g.N)

        def paint(comp0, attr, inverse):
            out = []
            for d in range(3):
                xs, ys, zs = self._coords(comp0 + d)
                X, Y, Z = np.meshgrid(xs, ys, zs, indexing="ij", sparse=True)
                a = np.full(shape, T(1) if inverse else T(0), dtype=T)
                for obj in reversed(self.geometry):  # earlier objects take priority (Geometry.jl:241)
                    m = obj.shape.contains(X, Y, Z)
                    v = getattr(obj.material, attr)
                    # get_perm_inv = one(T) / T(perm), get_sigma = T(sigma) (Geometry.jl:62-79)
                    a[np.broadcast_to(m, shape)] = (T(1) / T(v)) if inverse else T(v)
                out.append(a)
            return out

        if need_eps:
            arrays["eps_inv"] = paint(EX, "epsilon", True)
        if need_mu:
            arrays["mu_inv"] = paint(HX, "mu", True)
        if need_sd:
            arrays["sigma_D"] = paint(EX, "sigma_D", False)
        if need_sb:
            arrays["sigma_B"] = paint(HX, "sigma_B", False)
        # Kerr coefficient on the centre grid (Geometry.jl:610-635): later objects first, earlier ones
        # paint over them, objects without chi3 paint nothing
        if any(o.material.chi3 is not None for o in self.geometry):
            xs, ys, zs = self._coords(CENTER)
            X, Y, Z = np.meshgrid(xs, ys, zs, indexing="ij", sparse=True)
            a = np.zeros(shape, dtype=T)
            for obj in reversed(self.geometry):
                v = T(0) if obj.material.chi3 is None else T(obj.material.chi3)
                if v != 0:
                    a[np.broadcast_to(obj.shape.contains(X, Y, Z), shape)] = v
            arrays["chi3"] = a
        # dispersive poles: sigma rasterised on the Ex grid, shared by x/y/z (Geometry.jl:1146-1177)
        uniq = []
        for obj in self.geometry:
            for s in obj.material.susceptibilities:
                key = (s.omega_0, s.gamma)
                if key not in uniq:
                    uniq.append(key)
        xs, ys, zs = self._coords(EX)
        X, Y, Z = np.meshgrid(xs, ys, zs, indexing="ij", sparse=True)
        for key in uniq:
            # _rasterize_pole_sigma! (Geometry.jl:1082-1125): last object first; an object that does not
            # carry this pole is skipped (it does NOT clear what a lower-priority object painted); the
            # first matching susceptibility of an object wins
            sg = np.zeros(shape, dtype=T)
            for obj in reversed(self.geometry):
                match = [s for s in obj.material.susceptibilities if (s.omega_0, s.gamma) == key]
                if not match:
                    continue
                sg[np.broadcast_to(obj.shape.contains(X, Y, Z), shape)] = T(match[0].sigma)
            poles.append((key[0], key[1], sg))
        return arrays, poles

    def _zero_pole_sigma_in_pml(self, sigma):
        """Geometry.jl:1180-1232: dispersive sigma is removed inside the PML."""
        if self.boundaries is None:
            return sigma
        g = self.grid
        xs = self._coords(EX)
        mask = np.zeros(sigma.shape, dtype=bool)
        for a in range(3):
            half = np.asarray(g.cell_size[a] / g.T(2))  # cell_size ./ 2 stays T
            lo = g.cell_center[a] - float(half)
            hi = g.cell_center[a] + float(half)
            pl, pr = g.T(self.boundaries[a][0]), g.T(self.boundaries[a][1])
            m1 = np.zeros(g.N[a], dtype=bool)
            if pl > 0:
                m1 |= xs[a] < lo + float(pl)
            if pr > 0:
                m1 |= xs[a] > hi - float(pr)
            shp = [1, 1, 1]
            shp[a] = g.N[a]
            mask |= m1.reshape(shp)
        out = sigma.copy()
        out[mask] = 0
        return out

    def _apply_absorbers(self, arrays):
        """Geometry.jl:708-789 _apply_absorbers!: additive sigma ramp on the outer layers."""
        if self.absorbers is None:
            return arrays
        g, T = self.grid, self.T
        any_abs = any(ax is not None and any(s is not None for s in ax) for ax in self.absorbers)
        if not any_abs:
            return arrays
        for key in ("sigma_D", "sigma_B"):
            if arrays[key] is None:
                arrays[key] = [np.zeros(tuple(g.N), dtype=T) for _ in range(3)]
        for axis, ax in enumerate(self.absorbers):
            if ax is None:
                continue
            for side, ab in enumerate(ax):
                if ab is None:
                    continue
                L = T(ab.num_layers) * g.dl[axis]
                if ab.sigma_max > 0:
                    smax = T(ab.sigma_max)
                else:
                    smax = T(-(ab.sigma_order + 1) * math.log(1e-6) / (2.0 * float(L)))
                for key, comp0 in (("sigma_D", EX), ("sigma_B", HX)):
                    for c in range(3):
                        st = component_stagger(comp0 + c)
                        n_axis = g.N[axis] + st[axis]
                        for layer in range(1, min(ab.num_layers, n_axis) + 1):
                            dn = (ab.num_layers - layer + 1) / ab.num_layers
                            val = T(float(smax) * dn ** ab.sigma_order)
                            idx = (n_axis - layer + 1) if side == 1 else layer
                            if idx > g.N[axis]:
                                continue  # staggered extra cell: never read by the kernels
                            sl = [slice(None)] * 3
                            sl[axis] = idx - 1
                            arrays[key][c][tuple(sl)] += val
        return arrays

    # ------------------------------------------------------------ prepare
    def host_prepare(self):
        """Host half of prepare_simulation! (Simulation.jl:92-283): everything the hot path is
        handed — sigma profiles, material / pole arrays, source boxes + amplitudes, monitor
        index boxes — with no device involved.  Idempotent."""
        if getattr(self, "_host_ready", False):
            return
        g, T = self.grid, self.T
        self.slabs = chunking.z_slab_partition(g, self.boundaries, self.nranks, rule=self.slab_rule)
        self.z_start, self.nz_local = self.slabs[self.rank]
        # boundaries (Boundaries.jl:99-164): sigma_B and sigma_D profiles are identical
        # sigma profiles per field group [H, E].  Periodic / PEC / PMC sides drop their PML
        # (eff_boundaries, Boundaries.jl:100-110); reference quirk: sigma_Dz is built from the
        # raw sim.boundaries[3] (Boundaries.jl:154-161)
        self.sigma = None
        if self.boundaries is not None:
            def eff(a, raw):
                b = list(self.boundaries[a])
                if self.bc_codes is not None and not raw:
                    b = [0.0 if self.bc_codes[2 * a + sd] != 0 else b[sd] for sd in range(2)]
                return b
            self.sigma = [[g.compute_sigma(a, *eff(a, raw=(grp == 1 and a == 2))) for a in range(3)] for grp in range(2)]
        # geometry (Geometry.jl:450-663) + absorbers + poles
        arrays, poles = self._rasterize()
        if self.rasterizer == "device":
            if not self.geometry:
                raise ValueError("rasterizer='device' needs a geometry")
            if poles or self.user_poles or self.absorbers is not None or any(v is not None for v in self.user_arrays.values()):
                raise _lib.KhronosError("rasterizer='device' does not combine with dispersive poles, absorbers or "
                                        "user-supplied material arrays (those are prepared on the host)")
            for k in ("eps_inv", "mu_inv", "sigma_D", "sigma_B"):
                arrays[k] = None          # produced on the device in prepare_simulation
        for k, v in self.user_arrays.items():
            if v is not None:
                arrays[k] = [np.array(x, dtype=T) for x in v]
        if self.user_chi3 is not None:
            arrays["chi3"] = np.array(self.user_chi3, dtype=T)
        arrays = self._apply_absorbers(arrays)
        poles = poles + [(w, gam, np.asarray(s, dtype=T)) for (w, gam, s) in self.user_poles]
        poles = [(w, gam, self._zero_pole_sigma_in_pml(s)) for (w, gam, s) in poles]
        if poles:
            # chi1 semi-implicit correction folded into eps_inv (Geometry.jl:1236-1353)
            chi1 = np.zeros(tuple(g.N), dtype=T)
            for (w0, gam, s) in poles:
                dtd = float(g.dt)
                g1i = 1.0 / (1.0 + gam * math.pi * dtd)
                if w0 == 0.0:
                    c = T(g1i * (gam * (2 * math.pi) * dtd * dtd) / 2)
                else:
                    ww = (2 * math.pi) * w0 * dtd
                    c = T(g1i * (ww * ww) / 2)
                chi1 = chi1 + s * c
            if np.max(np.abs(chi1)) > 0:
                if arrays["eps_inv"] is None:
                    arrays["eps_inv"] = [np.full(tuple(g.N), T(1), dtype=T) for _ in range(3)]
                nz = chi1 != 0
                for d in range(3):
                    e = arrays["eps_inv"][d].copy()
                    e[nz] = e[nz] / (T(1) + e[nz] * chi1[nz])
                    arrays["eps_inv"][d] = e
        self.material_arrays = arrays
        self.poles = poles
        # sources (Sources.jl:40-135 assemble_sources)
        ct = np.complex64 if T is np.float32 else np.complex128
        self.source_data = []
        for src in self.sources:
            for comp in src.components:
                start, end = g.grid_volume(src.center, src.size, comp)
                dims = [end[a] - start[a] + 1 for a in range(3)]
                amp = self._source_amplitude(src, comp, start, dims).astype(ct)  # complex_backend_number.(...)
                self.source_data.append(dict(src=src, comp=comp, start=start, dims=dims, amp=amp))
        # monitors: auto-decimation (Monitors.jl:33-78), index boxes (:202-272)
        self.dft_monitors = []
        for m in self.monitors:
            if isinstance(m, FluxMonitor):
                self.dft_monitors.extend(m.monitors)
            else:
                self.dft_monitors.append(m)
        self._auto_decimate()
        for m in self.dft_monitors:
            m.start, m.end = g.grid_volume(m.center, m.size, m.component)
        self._host_ready = True

    def prepare_simulation(self, comm_id=None):
        """Simulation.jl:92-283 prepare_simulation!: host plan, then registration with the library."""
        if self.is_prepared:
            return
        self.host_prepare()
        L = _lib.lib()
        g, T = self.grid, self.T
        z_start, nzl = self.z_start, self.nz_local
        desc = _lib.GridDesc()
        desc.dtype = _lib.KHR_F32 if T is np.float32 else _lib.KHR_F64
        for a in range(3):
            desc.n[a] = g.N[a]
            desc.dl[a] = float(g.dl[a])
        desc.dt = float(g.dt)
        desc.z_start, desc.nz_local, desc.rank, desc.nranks = z_start, nzl, self.rank, self.nranks
        ctx = C.c_void_p()
        _lib.check(L.khr_ctx_create(self.device, C.byref(desc), C.byref(ctx)))
        self.ctx = ctx
        if self.complex_fields:
            _lib.check(L.khr_set_complex_fields(ctx))
        for a in range(3):
            if g.dlv[a] is not None:
                v = np.ascontiguousarray(g.dlv[a])
                _lib.check(L.khr_set_grid_spacing(ctx, a, v.ctypes.data, v.size))
        zsl = slice(z_start - 1, z_start - 1 + nzl)
        if self.sigma is not None:
            for grp in (_lib.GROUP_H, _lib.GROUP_E):
                for a in range(3):
                    s = np.ascontiguousarray(self.sigma[grp][a])
                    _lib.check(L.khr_set_pml_sigma(ctx, grp, a, s.ctypes.data, s.size))
        if self.rasterizer == "device":
            self.smoothed_voxels = self._device_rasterize(ctx)
        arrays = self.material_arrays
        kinds = {"eps_inv": _lib.MAT_EPS_INV, "mu_inv": _lib.MAT_MU_INV, "sigma_D": _lib.MAT_SIGMA_D,
                 "sigma_B": _lib.MAT_SIGMA_B}
        for k, kind in kinds.items():
            if arrays[k] is None:
                continue
            for d in range(3):
                a = np.asfortranarray(arrays[k][d][:, :, zsl])
                _lib.check(L.khr_set_material_array(ctx, kind, d, a.ctypes.data))
        if arrays.get("chi3") is not None:
            a = np.asfortranarray(arrays["chi3"][:, :, zsl])
            _lib.check(L.khr_set_material_array(ctx, _lib.MAT_CHI3, 0, a.ctypes.data))
        for (w0, gam, s) in self.poles:
            a = np.asfortranarray(s[:, :, zsl])
            pid = C.c_int32()
            _lib.check(L.khr_pole_register(ctx, w0, gam, a.ctypes.data, C.byref(pid)))
        self.source_ids = []
        i3 = C.c_int32 * 3
        for sd in self.source_data:
            amp = sd["amp"]
            ampT = np.empty(amp.shape + (2,), dtype=T)
            ampT[..., 0] = amp.real
            ampT[..., 1] = amp.imag
            buf = np.ascontiguousarray(np.transpose(ampT, (2, 1, 0, 3)))  # x fastest, (re,im) innermost
            tpf = sd["src"].time_profile
            tp = (C.c_double * 4)(*tpf.params(T))
            sid = C.c_int32()
            _lib.check(L.khr_source_register(ctx, sd["comp"], i3(*sd["start"]), i3(*sd["dims"]), buf.ctypes.data,
                                             tpf.kind, tp, C.byref(sid)))
            self.source_ids.append((sid.value, sd["src"], sd["comp"], sd["start"], sd["dims"]))
        for m in self.dft_monitors:
            fr = (C.c_double * len(m.frequencies))(*[float(T(f)) for f in m.frequencies])
            mid = C.c_int32()
            _lib.check(L.khr_monitor_register(ctx, m.component, i3(*m.start), i3(*m.end), len(m.frequencies), fr,
                                              m.decimation, C.byref(mid)))
            m.id = mid.value
        for a in range(3):
            if self.periodic[a] and self.complex_fields:
                # phase = exp(i * bloch_k * L), L = sim.cell_size[axis] (Chunking.jl:1745-1747)
                _lib.check(L.khr_set_bloch(ctx, a, float(self.bloch_k[a]) * float(g.cell_size[a])))
            elif self.periodic[a]:
                _lib.check(L.khr_set_periodic(ctx, a, 1))
        _lib.check(L.khr_finalize_plan(ctx))
        if self.nranks > 1:
            if comm_id is None:
                raise _lib.KhronosError("nranks > 1 needs the 128-byte NCCL unique id (see distributed.py)")
            buf = (C.c_char * 128).from_buffer_copy(bytes(comm_id))
            _lib.check(L.khr_comm_init(ctx, buf, self.nranks, self.rank))
        self.is_prepared = True

    def geometry_objects(self):
        """(KhrObject array, kinds mask) for khr_geometry_rasterize: get_perm_inv / get_sigma of every
        object (Geometry.jl:64-81), needs_perm / needs_conductivities over the scene (:317-352)."""
        T = self.T
        objs = (_lib.KhrObject * len(self.geometry))()
        for q, ob in enumerate(self.geometry):
            o, m = objs[q], ob.material
            if isinstance(ob.shape, Ball):
                o.kind = _lib.SHAPE_SPHERE
                o.size[0] = ob.shape.radius
            elif isinstance(ob.shape, Cuboid):
                o.kind = _lib.SHAPE_CUBOID
                for k in range(3):
                    o.size[k] = ob.shape.size[k]
                if ob.shape.axes is not None:
                    for k in range(9):
                        o.axes[k] = float(ob.shape.axes.reshape(9)[k])
            elif isinstance(ob.shape, Cylinder):
                o.kind = _lib.SHAPE_CYLINDER
                o.size[0], o.size[1] = ob.shape.radius, ob.shape.height
                for k in range(3):
                    o.axes[k] = float(ob.shape.axis[k])
            else:
                raise _lib.KhronosError("the device rasterizer knows Ball, Cuboid and Cylinder shapes")
            for k in range(3):
                o.center[k] = ob.shape.center[k]
                o.eps_inv[k] = float(T(1) / T(m.epsilon))     # one(T) / T(perm)
                o.mu_inv[k] = float(T(1) / T(m.mu))
                o.sigma_d[k] = float(T(m.sigma_D))
                o.sigma_b[k] = float(T(m.sigma_B))
        mask = 0
        if any(o.material.epsilon != 1.0 for o in self.geometry):
            mask |= 1 << _lib.MAT_EPS_INV
        if any(o.material.mu != 1.0 for o in self.geometry):
            mask |= 1 << _lib.MAT_MU_INV
        if any(o.material.sigma_D != 0.0 for o in self.geometry):
            mask |= 1 << _lib.MAT_SIGMA_D
        if any(o.material.sigma_B != 0.0 for o in self.geometry):
            mask |= 1 << _lib.MAT_SIGMA_B
        return objs, mask

    def component_origins(self):
        return np.ascontiguousarray([self.grid.component_origin(c) for c in (EX, EY, EZ, HX, HY, HZ)], dtype=np.float64)

    def _device_rasterize(self, ctx):
        objs, mask = self.geometry_objects()
        if mask == 0:
            return [0, 0, 0]
        org = self.component_origins().reshape(18)
        mode = {None: 0, "volume": 1, "anisotropic": 2}[self.subpixel_smoothing]
        cnt = (C.c_int64 * 3)()
        _lib.check(_lib.lib().khr_geometry_rasterize(ctx, objs, len(self.geometry), mask, mode,
                                                     org.ctypes.data_as(C.POINTER(C.c_double)), cnt))
        return list(cnt)

    def get_material(self, kind, comp):
        """Dense (Nx,Ny,Nz_local) copy of a device material array (kind: 'eps_inv', 'mu_inv', 'sigma_D', 'sigma_B')."""
        k = {"eps_inv": _lib.MAT_EPS_INV, "mu_inv": _lib.MAT_MU_INV, "sigma_D": _lib.MAT_SIGMA_D,
             "sigma_B": _lib.MAT_SIGMA_B, "chi3": _lib.MAT_CHI3}[kind]
        out = np.empty((self.Nx, self.Ny, self.nz_local), dtype=self.T, order="F")
        _lib.check(_lib.lib().khr_material_read(self.ctx, k, comp, out.ctypes.data))
        return out

    def _source_amplitude(self, src, comp, start, dims):
        """Sources.jl:107-135 _fill_amplitude_data!: weight * amplitude * profile (Complex{Float64})."""
        g = self.grid
        origin = g.component_origin(comp)
        d = [float(v) for v in g.dl]
        lo = [c - s / 2 for c, s in zip(src.center, src.size)]
        hi = [c + s / 2 for c, s in zip(src.center, src.size)]
        amp = np.zeros(tuple(dims), dtype=np.complex128)
        # the weight is separable per axis (product over dims in utils.jl:494)
        w = []
        for a in range(3):
            wa = np.zeros(dims[a])
            for i in range(1, dims[a] + 1):
                p = origin[a] + (i + start[a] - 2) * d[a]
                wa[i - 1] = interpolation_weight([p], [lo[a]], [hi[a]], [src.size[a]], 1, [d[a]])
            w.append(wa)
        # weight = ((1.0 * wx) * wy) * wz exactly as the reference's running product
        wgt = ((1.0 * w[0])[:, None, None] * w[1][None, :, None]) * w[2][None, None, :]
        if src.profile is None:
            amp[...] = wgt * src.amplitude * 1.0
        else:
            coords = [origin[a] + (np.arange(dims[a]) + start[a] - 1) * d[a] for a in range(3)]
            X, Y, Z = np.meshgrid(*coords, indexing="ij", sparse=True)
            prof = src.profile([X, Y, Z], comp)
            amp[...] = wgt * src.amplitude * prof
        return amp

    def _auto_decimate(self):
        """Monitors.jl:33-78 auto_decimate!: monitors the user left at decimation 1 get D_max."""
        for m in self.dft_monitors:
            m.decimation = m.user_decimation
        f_max = 0.0
        for s in self.sources:
            f_max = max(f_max, s.time_profile.f_max(self.T))
        if f_max <= 0:
            return
        d_max = max(1, int(math.floor(1.0 / (2.0 * f_max * float(self.grid.dt)))))
        if d_max <= 1:
            return
        for m in self.dft_monitors:
            if m.user_decimation == 1:
                m.decimation = d_max

    # --------------------------------------------------------------- step
    def _push_host_amplitudes(self):
        L = _lib.lib()
        T = self.T
        for sid, src, comp, _, _ in self.source_ids:
            if src.time_profile.kind == _lib.TIME_HOST:
                t = float(T(self.timestep) * self.grid.dt)
                if comp < 3:
                    t = t + float(self.grid.dt / T(2))
                a = complex(src.time_profile.src_func(T(t)))
                _lib.check(L.khr_source_set_amplitude(self.ctx, sid, a.real, a.imag))

    def step(self, n=1):
        """Kernels.jl:20-88 step! (n of them, the loop runs inside the library)."""
        if not self.is_prepared:
            self.prepare_simulation()
        L = _lib.lib()
        if any(s.time_profile.kind == _lib.TIME_HOST for s in self.sources):
            for _ in range(n):
                self._push_host_amplitudes()
                _lib.check(L.khr_step(self.ctx, 1))
                self.timestep += 1
        else:
            _lib.check(L.khr_step(self.ctx, int(n)))
            self.timestep += int(n)

    def round_time(self):
        """Simulation.jl:22."""
        return float(self.T(self.timestep) * self.grid.dt)

    def run(self, until=None, until_after_sources=None):
        """Simulation.jl:295-382 run(sim; until | until_after_sources).

        `until` is an absolute simulation time: stepping stops at the first step with
        round_time(sim) > until (run_until, Simulation.jl:384-387).  With
        `until_after_sources` the loop first runs while round_time <= last_source_time
        and then applies the same predicate to the given number (reference behaviour).
        Either argument may be a callable `f(sim) -> bool` (stop predicate).
        """
        if until is None and until_after_sources is None:
            raise ValueError("Must specify a terminating conditon in the run functions.")
        if until is not None and until_after_sources is not None:
            raise ValueError("Must specify a single terminating conditon in the run functions.")
        if not self.is_prepared:
            self.prepare_simulation()
        T, dt = self.T, self.grid.dt
        start = self.timestep

        def steps_while(cond, n0):
            n = n0
            while cond(float(T(n) * dt)):
                n += 1
            return n

        n = self.timestep
        if until_after_sources is not None:
            cut = [s.time_profile.cutoff() for s in self.sources]
            if any(c is None for c in cut):
                raise ValueError("Cutoff not possible with CW source...")
            last = float(max(T(c) for c in cut))
            n = steps_while(lambda t: t <= last, n)
            self.step(n - self.timestep)
        stop = until if until is not None else until_after_sources
        if callable(stop):
            while not stop(self):
                self.step(1)
        else:
            n = steps_while(lambda t: not (t > float(stop)), self.timestep)
            self.step(n - self.timestep)
        self.sync()
        return self.timestep - start

    def run_benchmark(self, n=110):
        """Simulation.jl:492-527: n steps, clock restarted after 10, Mcells/s returned."""
        if not self.is_prepared:
            self.prepare_simulation()
        self.step(10)
        self.sync()
        self.step(n - 10)
        ms = self.last_step_ms()
        return self.Nx * self.Ny * self.Nz * (n - 10) / (ms * 1e-3) / 1e6

    def sync(self):
        _lib.check(_lib.lib().khr_sync(self.ctx))

    def last_step_ms(self):
        ms = C.c_double()
        nl = C.c_int64()
        _lib.check(_lib.lib().khr_last_step_timing(self.ctx, C.byref(ms), C.byref(nl)))
        self.last_launches = nl.value
        return ms.value

    def reset_fields(self):
        """Simulation.jl:571-637 reset_fields!."""
        _lib.check(_lib.lib().khr_reset_fields(self.ctx))
        self.timestep = 0

    # -------------------------------------------------------------- access
    def get_field(self, comp, part="real"):
        """Local slab of a field component, dense (Nx,Ny,Nz_local) (Visualization.jl:294-333);
        part="imag" / "complex" for complex fields (Bloch boundaries)."""
        L = _lib.lib()
        out = np.empty((self.Nx, self.Ny, self.nz_local), dtype=self.T, order="F")
        if part in ("real", "complex"):
            _lib.check(L.khr_field_read(self.ctx, comp, out.ctypes.data))
        if part == "real":
            return out
        im = np.empty((self.Nx, self.Ny, self.nz_local), dtype=self.T, order="F")
        _lib.check(L.khr_field_read_imag(self.ctx, comp, im.ctypes.data))
        return im if part == "imag" else out + 1j * im

    def set_field(self, comp, arr):
        a = np.asfortranarray(np.asarray(arr, dtype=self.T))
        if a.shape != (self.Nx, self.Ny, self.nz_local):
            raise ValueError("field shape must be (Nx,Ny,Nz_local)")
        _lib.check(_lib.lib().khr_field_write(self.ctx, comp, a.ctypes.data))

    def get_dft(self, monitor):
        """Array(md.fields): complex (nx,ny,nz,nf) (FluxMonitor.jl:99-102), Complex{T}, zero-copy view
        of the buffer the device->host copy filled."""
        n = [monitor.end[a] - monitor.start[a] + 1 for a in range(3)] + [len(monitor.frequencies)]
        ct = np.complex64 if self.T is np.float32 else np.complex128
        raw = np.empty(tuple(n[::-1]), dtype=ct)  # memory order: (re,im) pairs, x fastest, f slowest
        _lib.check(_lib.lib().khr_monitor_read(self.ctx, monitor.id, raw.ctypes.data))
        return raw.transpose(3, 2, 1, 0)

    def get_flux(self, fm, dft=None):
        """FluxMonitor.jl:92-156 get_flux: Σ Re(E1·conj(H2) − E2·conj(H1))·dA per frequency.

        Default: reduced on the device (`khr_flux`), only the nf values come back.  `dft` may
        carry the four (already rank-reduced, `distributed.reduce_dft`) DFT arrays of a box that
        is split across ranks; those go through the same formula on the host.
        """
        if dft is None:
            ids = (C.c_int32 * 4)(*[m.id for m in fm.monitors])
            nf = len(fm.frequencies)
            out = (C.c_double * nf)()
            _lib.check(_lib.lib().khr_flux(self.ctx, ids, fm.normal, out, nf))
            return np.array(list(out), dtype=np.float64)
        T = self.T
        ct = np.complex64 if T is np.float32 else np.complex128
        arrs = dft if dft is not None else [self.get_dft(m) for m in fm.monitors]
        arrs = [np.asarray(a).astype(ct) for a in arrs]
        nrm, (t1, t2) = fm.normal, fm.tangential

        def avg(a):  # _avg_dim: mean of the two planes when the box is 2 cells thick
            if a.shape[nrm] >= 2:
                idx0 = [slice(None)] * 4
                idx1 = [slice(None)] * 4
                idx0[nrm], idx1[nrm] = 0, 1
                return (a[tuple(idx0)] + a[tuple(idx1)]) / T(2)
            return np.take(a, 0, axis=nrm)

        e1, e2, h1, h2 = [avg(a) for a in arrs]  # each (n_t1, n_t2, nf)
        n1 = min(a.shape[0] for a in (e1, e2, h1, h2))
        n2 = min(a.shape[1] for a in (e1, e2, h1, h2))
        e1, e2, h1, h2 = [a[:n1, :n2, :] for a in (e1, e2, h1, h2)]
        dA = float(self.grid.dl[t1]) * float(self.grid.dl[t2])
        re1 = e1.real * h2.real + e1.imag * h2.imag
        re2 = e2.real * h1.real + e2.imag * h1.imag
        s = (re1 - re2).astype(np.float64) * dA
        return s.sum(axis=(0, 1))

    def _plane_bases(self, fm):
        """md.e1_base, e2_base, h1_base, h2_base (Monitors.jl:398-420): physical position of
        dft[1,1,1] = component origin + (start_idx - 1) * Δ, Float64."""
        g = self.grid
        out = []
        for m in fm.monitors:
            o = g.component_origin(m.component)
            out.append([o[a] + (float(m.start[a]) - 1.0) * float(g.dl[a]) for a in range(3)])
        return np.asarray(out, dtype=np.float64)

    def far_field_points(self, fm):
        """Observation points of compute_far_field (Near2Far.jl:1004-1021): theta/phi/r grid (phi
        outer, theta inner) or the explicit list."""
        if fm.theta is not None and fm.phi is not None:
            pts = [[fm.r * math.sin(t) * math.cos(p), fm.r * math.sin(t) * math.sin(p), fm.r * math.cos(t)]
                   for p in fm.phi for t in fm.theta]
            return np.asarray(pts, dtype=np.float64)
        if fm.observation_points is not None:
            return np.asarray(fm.observation_points, dtype=np.float64).reshape(-1, 3)
        raise ValueError("Near2FarMonitorData must have either theta/phi or observation_points")

    def compute_far_field(self, fm, obs_points=None):
        """Near2Far.jl:998-1031 compute_far_field (free-space Green's function; layer stacks are not
        on this path): EH complex128 (nobs, 6, nf), evaluated on the device (`khr_near2far`)."""
        obs = np.ascontiguousarray(self.far_field_points(fm) if obs_points is None
                                   else np.asarray(obs_points, dtype=np.float64).reshape(-1, 3))
        nobs, nf = obs.shape[0], len(fm.frequencies)
        ids = (C.c_int32 * 4)(*[m.id for m in fm.monitors])
        bases = np.ascontiguousarray(self._plane_bases(fm).reshape(12))
        freqs = np.asarray([float(f) for f in fm.frequencies], dtype=np.float64)  # Float64.(monitor.frequencies)
        out = np.zeros(2 * nobs * 6 * nf, dtype=np.float64)
        dp = C.POINTER(C.c_double)
        _lib.check(_lib.lib().khr_near2far(self.ctx, ids, fm.normal, fm.normal_sign, fm.medium_eps, fm.medium_mu,
                                           bases.ctypes.data_as(dp), freqs.ctypes.data_as(dp), nf,
                                           obs.ctypes.data_as(dp), nobs, out.ctypes.data_as(dp)))
        z = out[0::2] + 1j * out[1::2]
        return z.reshape(nf, 6, nobs).transpose(2, 1, 0)

    @staticmethod
    def compute_far_field_power(EH, theta, phi, eps=1.0, mu=1.0):
        """Near2Far.jl:1043-1072: radiated power per solid angle from the first frequency of EH."""
        Z = math.sqrt(mu / eps)
        power = np.zeros((len(theta), len(phi)))
        idx = 0
        for ip, ph in enumerate(phi):
            for it, th in enumerate(theta):
                ex, ey, ez = EH[idx, 0, 0], EH[idx, 1, 0], EH[idx, 2, 0]
                th_hat = (math.cos(th) * math.cos(ph), math.cos(th) * math.sin(ph), -math.sin(th))
                ph_hat = (-math.sin(ph), math.cos(ph), 0.0)
                e_th = ex * th_hat[0] + ey * th_hat[1] + ez * th_hat[2]
                e_ph = ex * ph_hat[0] + ey * ph_hat[1] + ez * ph_hat[2]
                power[it, ip] = 0.5 * (abs(e_th) ** 2 + abs(e_ph) ** 2) / Z
                idx += 1
        return power

    def get_diffraction_efficiencies(self, fm, max_order=5, k_inc=(0.0, 0.0, 0.0)):
        """DiffractionMonitor.jl:87-165: {(m, n): power per frequency} for the orders that propagate at
        some frequency (evanescent entries stay 0, as in the reference), evaluated on the device."""
        t1, t2 = fm.tangential
        nf, nord = len(fm.frequencies), 2 * int(max_order) + 1
        ids = (C.c_int32 * 4)(*[m.id for m in fm.monitors])
        freqs = np.asarray([float(f) for f in fm.frequencies], dtype=np.float64)
        power = np.zeros(nf * nord * nord, dtype=np.float64)
        prop = np.zeros(nf * nord * nord, dtype=np.int32)
        L = [float(v) for v in self.grid.cell_size]
        _lib.check(_lib.lib().khr_diffraction(self.ctx, ids, fm.normal, int(max_order), L[t1], L[t2], float(k_inc[t1]),
                                              float(k_inc[t2]), freqs.ctypes.data_as(C.POINTER(C.c_double)), nf,
                                              power.ctypes.data_as(C.POINTER(C.c_double)),
                                              prop.ctypes.data_as(C.POINTER(C.c_int32))))
        power, prop = power.reshape(nf, nord, nord), prop.reshape(nf, nord, nord)
        out = {}
        for im in range(nord):
            for i_n in range(nord):
                if prop[:, im, i_n].any():
                    out[(im - max_order, i_n - max_order)] = power[:, im, i_n].copy()
        return out

    @staticmethod
    def compute_LEE(power, theta, phi, cone_half_angle=math.pi / 2):
        """Near2Far.jl:1082-1124 compute_LEE: power inside a cone over the hemispherical power,
        trapezoidal weights with the sin(theta) Jacobian."""
        power = np.asarray(power, dtype=np.float64)
        nt, npz = len(theta), len(phi)
        dth = np.diff(theta) if nt > 1 else np.array([0.0])
        dph = np.diff(phi) if npz > 1 else np.array([0.0])
        p_total = p_cone = 0.0
        for ip in range(npz):
            if ip == 0:
                wp = dph[0] / 2 if len(dph) > 0 else 2 * math.pi
            elif ip == npz - 1:
                wp = dph[-1] / 2
            else:
                wp = (dph[ip - 1] + dph[ip]) / 2
            for it in range(nt):
                if it == 0:
                    wt = dth[0] / 2 if len(dth) > 0 else math.pi / 2
                elif it == nt - 1:
                    wt = dth[-1] / 2
                else:
                    wt = (dth[it - 1] + dth[it]) / 2
                integrand = power[it, ip] * math.sin(theta[it]) * wt * wp
                if theta[it] <= math.pi / 2:
                    p_total += integrand
                if theta[it] <= cone_half_angle:
                    p_cone += integrand
        return p_cone / p_total if p_total > 0 else 0.0

    def compute_mode_amplitudes(self, fm, mode_fields):
        """ModeMonitor.jl:345-515 compute_mode_amplitudes with the surface sums on the device
        (`khr_mode_overlap`).  mode_fields: complex (4, n1, n2, nf) = mode e1, e2, h1, h2 already
        interpolated onto the DFT grid (:427-457).  Returns (a_plus, a_minus)."""
        m = np.asarray(mode_fields, dtype=np.complex128)
        if m.ndim != 4 or m.shape[0] != 4 or m.shape[3] != len(fm.frequencies):
            raise ValueError("mode_fields must be complex (4, n1, n2, nf)")
        n1, n2, nf = m.shape[1], m.shape[2], m.shape[3]
        flat = np.ascontiguousarray(m.transpose(0, 3, 2, 1)).ravel()  # [4][nf][n2][n1]
        ri = np.empty(2 * flat.size, dtype=np.float64)
        ri[0::2] = flat.real
        ri[1::2] = flat.imag
        ids = (C.c_int32 * 4)(*[mm.id for mm in fm.monitors])
        out = np.zeros(5 * nf, dtype=np.float64)
        dp = C.POINTER(C.c_double)
        _lib.check(_lib.lib().khr_mode_overlap(self.ctx, ids, fm.normal, ri.ctypes.data_as(dp), n1, n2, nf,
                                               out.ctypes.data_as(dp)))
        o = out.reshape(nf, 5)
        a_plus = np.zeros(nf, dtype=np.complex128)
        a_minus = np.zeros(nf, dtype=np.complex128)
        for k in range(nf):
            P = o[k, 0]
            if abs(P) > 1e-30:
                a_plus[k] = complex(o[k, 1], o[k, 2]) / (4.0 * P)
                a_minus[k] = complex(o[k, 3], o[k, 4]) / (4.0 * P)
        return a_plus, a_minus

    def get_mode_transmission(self, fm, mode_fields):
        """ModeMonitor.jl:515-518: |a+|^2 per frequency."""
        return np.abs(self.compute_mode_amplitudes(fm, mode_fields)[0]) ** 2

    def set_profiling(self, mode):
        _lib.check(_lib.lib().khr_set_profiling(self.ctx, int(mode)))

    def kernel_stats(self):
        """Per-kernel live CUDA-event timing and bytes model (khr_kernel_stat_get)."""
        L = _lib.lib()
        n = C.c_int32()
        _lib.check(L.khr_kernel_stat_get(self.ctx, -1, None, C.byref(n)))
        out = []
        for i in range(n.value):
            st = _lib.KernelStat()
            _lib.check(L.khr_kernel_stat_get(self.ctx, i, C.byref(st), None))
            out.append(dict(name=st.name.decode(), launches=st.launches, total_ms=st.total_ms,
                            cells_per_launch=st.cells_per_launch, alg_bytes_per_launch=st.alg_bytes_per_launch,
                            ref_model_bytes_per_launch=st.ref_model_bytes_per_launch,
                            ctas=st.ctas, uniform_ctas=st.uniform_ctas))
        return out

    def graph_info(self):
        """(kernels in the CUDA graph of a step, graph replays so far); (0, 0) when no graph is in use."""
        k, r = C.c_int64(), C.c_int64()
        _lib.check(_lib.lib().khr_graph_info(self.ctx, C.byref(k), C.byref(r)))
        return k.value, r.value

    def comm_stats(self):
        """(ms the main stream waited for halo planes, number of exchanges) since the profiling reset."""
        ms, n = C.c_double(), C.c_int64()
        _lib.check(_lib.lib().khr_comm_stat_get(self.ctx, C.byref(ms), C.byref(n)))
        return ms.value, n.value

    def monitor_norm(self, monitor):
        v = C.c_double()
        _lib.check(_lib.lib().khr_monitor_norm(self.ctx, monitor.id, C.byref(v)))
        return v.value

    def monitor_norms(self):
        """dft_fields_norm of every DFT monitor (Simulation.jl:528-539), one call."""
        n = len(self.dft_monitors)
        out = (C.c_double * n)()
        _lib.check(_lib.lib().khr_monitor_norms(self.ctx, out, n))
        return list(out)

    def voxel_census(self):
        c = (C.c_int64 * 4)()
        _lib.check(_lib.lib().khr_voxel_census(self.ctx, c))
        return list(c)

    def device_bytes(self):
        b = C.c_int64()
        _lib.check(_lib.lib().khr_device_bytes(self.ctx, C.byref(b)))
        return b.value

    def close(self):
        if self.ctx is not None:
            _lib.lib().khr_ctx_destroy(self.ctx)
            self.ctx = None
            self.is_prepared = False

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def stop_when_dft_decayed(tolerance=1e-6, minimum_runtime=0.0, maximum_runtime=float("inf")):
    """Simulation.jl:411-485 stop_when_dft_decayed: a stop predicate for `run(until_after_sources=...)`.
    Each DFT monitor tracks the step-to-step change of its norm (dft_fields_norm, `khr_monitor_norms`:
    one device reduction for all stale monitors, cached between DFT updates) relative to the largest
    change it has seen; the run stops when every monitor that has seen signal reports
    rel_change <= tolerance.  Restated literally, including its behaviour with decimated monitors:
    on a step without a DFT update the norm does not change, so rel_change = 0 there."""
    if minimum_runtime > maximum_runtime:
        raise ValueError("Minimum runtime (%s) cannot be greater than maximum runtime (%s)." % (minimum_runtime, maximum_runtime))
    state = {}

    def _stop(sim):
        t = sim.round_time()
        if t < minimum_runtime:
            return False
        if t > maximum_runtime:
            return True
        all_converged, n_active = True, 0
        for i, current in enumerate(sim.monitor_norms()):
            if current == 0.0:
                continue
            if i not in state:
                state[i] = (current, 0.0)
                all_converged = False
                continue
            prev, maxchange = state[i]
            change = abs(current - prev)
            maxchange = max(maxchange, change)
            state[i] = (current, maxchange)
            if maxchange == 0.0:
                all_converged = False
                continue
            n_active += 1
            if change / maxchange > tolerance:
                all_converged = False
        if n_active == 0:
            return False
        return all_converged

    return _stop


def run(sim, until=None, until_after_sources=None):
    """Khronos.run(sim; until=..., until_after_sources=...)."""
    return sim.run(until=until, until_after_sources=until_after_sources)


def step(sim):
    """Khronos.step!(sim)."""
    sim.step(1)


def run_benchmark(sim, n=110):
    return sim.run_benchmark(n)
