"""GPU parity: the CUDA path (through the C ABI) against the CPU oracle on the same inputs.

Tolerances are the north star's: fields, DFT and flux within 1e-5 relative L2 for
Float32 and 1e-12 for Float64 after N steps.
"""
import numpy as np
import pytest

import khronos_b200 as kb
from common import Pair, rel_l2

pytestmark = pytest.mark.gpu

TOL = {np.float32: 1e-5, np.float64: 1e-12}
CW = kb.ContinuousWaveSource(fcen=1.0)


def _check(p, nsteps, tol=None):
    p.step(nsteps)
    tol = tol or TOL[p.dtype]
    err = p.total_field_error()
    per = p.field_errors()
    assert err < tol, (err, per)
    for km, om in zip(p.kmon, p.omon):
        a = p.k.get_dft(km)
        b = p.o.get_dft(om)
        assert a.shape == b.shape
        e = rel_l2(a, b)
        assert e < tol, ("dft", km.component, e)
    return err


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_dipole_vacuum_pml(dtype):
    """benchmark/dipole.jl at 40^3: Ez point dipole, CW, PML on all sides, + Ez DFT plane."""
    p = Pair([4, 4, 4], 10, [1.0, 1.0, 1.0], dtype, sources=[(kb.EZ, [0, 0, 0], [0, 0, 0], CW)],
             monitors=[(kb.EZ, [0, 0, 0], [2, 2, 0], [0.8, 1.0, 1.2], 1), (kb.HX, [0, 0.3, 0], [2, 0, 2], [1.0], 1)])
    _check(p, 80)


def test_no_pml_interior_only():
    p = Pair([3, 2.5, 2], 12, None, np.float32, sources=[(kb.EX, [0.1, 0, 0], [0, 0, 0], CW)])
    _check(p, 50)


def test_ragged_sizes_and_asymmetric_pml():
    """Nx, Ny, Nz not multiples of 4 / of the tile; different PML per side; Hy line source."""
    p = Pair([4.3, 3.7, 2.9], 10, [[0.5, 1.0], [0.0, 0.8], [0.7, 0.0]], np.float32,
             sources=[(kb.HY, [0.2, -0.1, 0.1], [1.0, 0, 0], CW), (kb.EZ, [0, 0, 0], [0, 0, 0],
                      kb.GaussianPulseSource(fcen=1.0, fwidth=0.4))],
             monitors=[(kb.EY, [0, 0, 0.2], [4.3, 3.7, 0], [1.0, 1.1], 1)])
    _check(p, 70)


def test_per_voxel_eps_and_mu():
    rng = np.random.default_rng(1234)
    N = (44, 36, 40)
    eps = [(1.0 / rng.uniform(1.0, 4.0, N)).astype(np.float32) for _ in range(3)]
    mu = [(1.0 / rng.uniform(1.0, 1.5, N)).astype(np.float32) for _ in range(3)]
    p = Pair([4.4, 3.6, 4.0], 10, [1.0, 1.0, 1.0], np.float32, eps_inv=eps, mu_inv=mu,
             sources=[(kb.EZ, [0, 0, 0], [0, 0, 0], CW), (kb.HX, [0.5, 0, 0], [0, 1.0, 1.0], CW)],
             monitors=[(kb.EZ, [0, 0, 0], [0, 2, 2], [1.0], 1)])
    _check(p, 60)


def test_gaussian_pulse_switch_off():
    """sources_active flips off after the cutoff (Kernels.jl:27-35)."""
    tp = kb.GaussianPulseSource(fcen=1.0, fwidth=2.0, cutoff_scale=3.0)
    p = Pair([3, 3, 3], 10, [0.5, 0.5, 0.5], np.float32, sources=[(kb.EY, [0, 0, 0], [0, 0, 0], tp)])
    nsteps = int(tp.cutoff() / float(p.grid.dt)) + 30
    _check(p, nsteps)


def test_absorber_and_conductivity():
    """material sigma_D/sigma_B (absorber ramp) with and without PML (Helpers.jl:39-69, 141-154)."""
    ab = [[kb.Absorber(8, 3), kb.Absorber(8, 3)], None, [None, kb.Absorber(6, 2)]]
    p = Pair([3, 3, 3], 10, [[0, 0], [0.6, 0.6], [0.5, 0]], np.float32, absorbers=ab,
             sources=[(kb.EZ, [0, 0, 0], [0, 0, 0], CW)])
    # (the sigma arrays handed to the oracle already carry the ramp; test_host_maps checks the ramp itself)
    _check(p, 60)


@pytest.mark.parametrize("kind", ["drude", "lorentz", "both"])
def test_dispersive_ade(kind):
    """uled-like: a Drude / Lorentz slab inside vacuum, ADE fused into the E half-step."""
    N = (40, 40, 30)
    sg = np.zeros(N, dtype=np.float32)
    sg[:, :, 8:14] = 1.0
    sg2 = np.zeros(N, dtype=np.float32)
    sg2[10:30, 10:30, 14:20] = 0.7
    poles = []
    if kind in ("drude", "both"):
        poles.append((0.0, 0.3, sg * 5.0))
    if kind in ("lorentz", "both"):
        poles.append((1.2, 0.1, sg2 if kind == "both" else sg))
    p = Pair([4, 4, 3], 10, [0.5, 0.5, 0.5], np.float32, poles=poles,
             sources=[(kb.EY, [0, 0, 0.9], [0, 0, 0], kb.ContinuousWaveSource(fcen=1.5))],
             monitors=[(kb.EY, [0, 0, -0.5], [2, 2, 0], [1.5], 1)])
    _check(p, 80)


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_six_poles_and_twelve_sources(dtype):
    """Any number of poles and sources, as the reference loops over all of them (Dispersive.jl:186-228,
    Sources.jl:330-340): 6 poles (3 Drude, 3 Lorentz, partly overlapping, one touching the PML) and 12 E-group
    + 9 H-group sources (points, lines, sheets; several hit the same tile and the same voxel) — more than the
    4 / 8 descriptors that ride in the kernel parameters, so the device tables are exercised."""
    N = (48, 40, 36)
    rng = np.random.default_rng(5)
    poles = []
    for q in range(6):
        sg = np.zeros(N, dtype=dtype)
        x0, y0, z0 = 6 + 5 * q, 4 + 3 * q, 8 + 3 * q
        sg[x0:x0 + 14, y0:y0 + 16, z0:z0 + 5] = dtype(0.4 + 0.3 * q)
        poles.append((0.0 if q % 2 == 0 else 0.9 + 0.2 * q, 0.1 + 0.05 * q, sg))
    srcs = []
    for q in range(12):
        comp = (kb.EX, kb.EY, kb.EZ)[q % 3]
        c = [float(v) for v in rng.uniform(-1.2, 1.2, 3)]
        size = [[0, 0, 0], [1.0, 0, 0], [0, 0.8, 0.6], [0.5, 0.5, 0]][q % 4]
        srcs.append((comp, c, size, kb.ContinuousWaveSource(fcen=0.8 + 0.05 * q)))
    for q in range(9):
        comp = (kb.HX, kb.HY, kb.HZ)[q % 3]
        c = [float(v) for v in rng.uniform(-1.0, 1.0, 3)]
        srcs.append((comp, c, [[0, 0, 0], [0, 0.7, 0]][q % 2], kb.GaussianPulseSource(fcen=1.0 + 0.05 * q, fwidth=0.5)))
    srcs.append((kb.EZ, srcs[0][1], [0, 0, 0], CW))      # a second source on the very same voxels as the first
    p = Pair([4.8, 4.0, 3.6], 10, [0.6, 0.6, 0.6], dtype, poles=poles, sources=srcs,
             monitors=[(kb.EY, [0, 0, 0.2], [3, 3, 0], [0.9, 1.1], 1), (kb.HZ, [0, 0.1, 0], [3, 0, 2.4], [1.0], 2)])
    _check(p, 70)


def test_random_state_single_step():
    """uniform [-1,1] initial fields (seed 1234), one step, no sources."""
    rng = np.random.default_rng(1234)
    p = Pair([3.2, 2.8, 3.0], 10, [0.5, 0.5, 0.5], np.float32)
    for comp in range(6):
        a = rng.uniform(-1, 1, tuple(p.grid.N)).astype(np.float32)
        p.o.set_field(comp, a)
        p.k.set_field(comp, a)
    _check(p, 1, tol=2e-6)
    _check(p, 4)


def test_flux_monitor_matches_oracle():
    fm = kb.FluxMonitor([0.5, 0, 0], [0, 2, 2], [0.9, 1.0, 1.1])
    mons = [(m.component, m.center, m.size, m.frequencies, 1) for m in fm.monitors]
    p = Pair([4, 4, 4], 10, [1.0, 1.0, 1.0], np.float32, sources=[(kb.EZ, [0, 0, 0], [0, 0, 0], CW)], monitors=mons)
    p.step(120)
    fm.monitors = p.kmon
    flux_gpu = p.k.get_flux(fm)
    flux_cpu = p.o.flux(0, p.omon)
    assert rel_l2(flux_gpu, flux_cpu) < 1e-5, (flux_gpu, flux_cpu)


def test_decimated_monitor_and_run_api():
    """auto-decimation (Monitors.jl:33-78) and run(until=...) stepping count."""
    p = Pair([3, 3, 3], 16, [0.5, 0.5, 0.5], np.float32, sources=[(kb.EZ, [0, 0, 0], [0, 0, 0], CW)],
             monitors=[(kb.EZ, [0, 0, 0], [1, 1, 0], [1.0], 1)])
    assert p.kmon[0].decimation == 16
    n = p.k.run(until=2.0)
    p.o.step(n)
    assert p.k.timestep == p.o.timestep
    assert rel_l2(p.k.get_dft(p.kmon[0]), p.o.get_dft(p.omon[0])) < 1e-5
    assert p.total_field_error() < 1e-5


def test_no_silent_fallback_and_errors():
    """bad arguments come back as explicit errors through the ABI."""
    import ctypes as C
    L = kb._lib.lib()
    d = kb._lib.GridDesc()
    d.dtype = 7
    for a in range(3):
        d.n[a] = 8
        d.dl[a] = 0.1
    d.dt, d.z_start, d.nz_local, d.rank, d.nranks = 0.05, 1, 8, 0, 1
    ctx = C.c_void_p()
    assert L.khr_ctx_create(0, C.byref(d), C.byref(ctx)) != 0
    assert b"dtype" in L.khr_last_error()
    d.dtype = 0
    assert L.khr_ctx_create(0, C.byref(d), C.byref(ctx)) == 0
    assert L.khr_step(ctx, 1) != 0  # finalize not called
    assert b"finalize" in L.khr_last_error()
    L.khr_ctx_destroy(ctx)


@pytest.mark.parametrize("normal", [0, 1, 2])
@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_device_flux_all_normals(normal, dtype):
    """khr_flux (FluxMonitor.jl:92-156 on the device) against the oracle's get_flux and against the
    same formula evaluated on the host from the four DFT arrays: the per-cell arithmetic is the
    reference's, only the Float64 summation order differs."""
    size = [2.0, 2.0, 2.0]
    size[normal] = 0.0
    center = [0.0, 0.0, 0.0]
    center[normal] = 0.5
    fm = kb.FluxMonitor(center, size, [0.9, 1.0, 1.1])
    mons = [(m.component, m.center, m.size, m.frequencies, 1) for m in fm.monitors]
    p = Pair([4, 4, 4], 10, [1.0, 1.0, 1.0], dtype, sources=[(kb.EZ, [0, 0, 0], [0, 0, 0], CW),
                                                              (kb.EX, [0.2, -0.1, 0.3], [0, 0, 0], CW)], monitors=mons)
    p.step(60)
    fm.monitors = p.kmon
    dev = p.k.get_flux(fm)
    host = p.k.get_flux(fm, dft=[p.k.get_dft(m) for m in fm.monitors])
    ora = np.asarray(p.o.flux(normal, p.omon))
    assert rel_l2(dev, host) < 1e-12, (dev, host)
    assert rel_l2(dev, ora) < TOL[dtype], (dev, ora)


def test_device_flux_argument_errors():
    fm = kb.FluxMonitor([0.5, 0, 0], [0, 2, 2], [1.0])
    mons = [(m.component, m.center, m.size, m.frequencies, 1) for m in fm.monitors]
    p = Pair([4, 4, 4], 10, [1.0, 1.0, 1.0], np.float32, sources=[(kb.EZ, [0, 0, 0], [0, 0, 0], CW)], monitors=mons)
    fm.monitors = list(reversed(p.kmon))       # H monitors first: rejected
    with pytest.raises(kb.KhronosError):
        p.k.get_flux(fm)
