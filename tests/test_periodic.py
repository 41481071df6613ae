"""Periodic boundary conditions (DataStructures.jl:150-161, Chunking.jl:1725-1770 wrap-around
connections, Boundaries.jl:100-110 eff_boundaries).

CPU part: the oracle's wrap is checked through an exact property — with every axis periodic
and no PML the update is invariant under cyclic translation, so moving the source by whole
cells must roll the fields bit for bit.  GPU part: the CUDA path against the oracle."""
import numpy as np
import pytest

import khronos_b200 as kb
from common import Pair, rel_l2

CW = kb.ContinuousWaveSource(fcen=1.0)
PER = [[kb.Periodic(), kb.Periodic()]] * 3


def _oracle_fields(center, nsteps, dtype=np.float32):
    from bridge import oracle_from_simulation
    # resolution 8: the cell size 0.125 and the shifts are exact in binary, so the source weights
    # of the two runs are bit-identical
    sim = kb.Simulation([2.0, 1.5, 1.0], [0, 0, 0], 8, [kb.UniformSource(CW, kb.EZ, center, [0, 0, 0]),
                                                         kb.UniformSource(CW, kb.HX, center, [0, 0, 0])],
                        boundaries=[[0.0, 0.0]] * 3, boundary_conditions=PER, dtype=dtype)
    o, _ = oracle_from_simulation(sim)
    o.step(nsteps)
    return [o.get_field(c) for c in range(6)]


def test_oracle_periodic_translation_invariance():
    a = _oracle_fields([0.0, 0.0, 0.0], 40)
    b = _oracle_fields([0.5, -0.375, 0.25], 40)       # +4, -3, +2 cells
    assert max(np.abs(f).max() for f in a) > 0
    for fa, fb in zip(a, b):
        assert np.array_equal(np.roll(fa, (4, -3, 2), axis=(0, 1, 2)), fb)
    # the wave has crossed the boundary: without the wrap the far corner would still be ~0
    assert np.abs(a[2][0, 0, 0]) > 0


def _oracle_bloch(center, nsteps, k, dtype=np.float64):
    from bridge import oracle_from_simulation
    bc = [[kb.Bloch(k[0]), kb.Bloch(k[0])], [kb.Bloch(k[1]), kb.Periodic()], [kb.Periodic(), kb.Bloch(k[2])]]
    sim = kb.Simulation([2.0, 1.5, 1.0], [0, 0, 0], 8, [kb.UniformSource(CW, kb.EZ, center, [0, 0, 0]),
                                                         kb.UniformSource(CW, kb.HX, center, [0, 0, 0])],
                        boundaries=[[0.0, 0.0]] * 3, boundary_conditions=bc, dtype=dtype)
    assert sim.complex_fields and sim.bloch_k == list(k)
    o, _ = oracle_from_simulation(sim)
    o.step(nsteps)
    return [o.get_field(c) + 1j * o.get_field(c, which="imag") for c in range(6)], sim


def test_oracle_bloch_k0_is_bit_identical_to_periodic():
    """Bloch(k = 0) allocates complex fields (Fields.jl:140-159) but its phase factor is exactly 1
    and is skipped (Chunking.jl:2165): real parts equal the Periodic run, imaginary parts stay 0."""
    a = _oracle_fields([0.25, 0.0, -0.125], 40, np.float64)
    b, _ = _oracle_bloch([0.25, 0.0, -0.125], 40, (0.0, 0.0, 0.0))
    for fa, fb in zip(a, b):
        assert np.array_equal(fa, fb.real) and not fb.imag.any()


def test_oracle_bloch_translation_covariance():
    """Exact symmetry of the discrete update with f(x + L) = f(x) exp(i k L) on every axis: moving the
    source by whole cells rolls the fields, and the cells that wrapped around pick up exp(-i k L)
    (positive shift) or exp(+i k L) (negative shift).  Pins the sign convention of both ghost copies
    (Chunking.jl:1735-1764) and the complex arithmetic of the wrap."""
    k = (0.7, -0.4, 1.1)
    a, sim = _oracle_bloch([0.0, 0.0, 0.0], 48, k)
    b, _ = _oracle_bloch([0.5, -0.375, 0.25], 48, k)       # +4, -3, +2 cells
    L = [float(v) for v in sim.grid.cell_size]
    shifts = (4, -3, 2)
    assert max(np.abs(f.imag).max() for f in a) > 1e-3      # the phase really mixes the two parts
    for fa, fb in zip(a, b):
        want = fa.copy()
        for ax, m in enumerate(shifts):
            want = np.roll(want, m, axis=ax)
            sl = [slice(None)] * 3
            if m > 0:
                sl[ax] = slice(0, m)
                want[tuple(sl)] *= np.exp(-1j * k[ax] * L[ax])
            else:
                sl[ax] = slice(m, None)
                want[tuple(sl)] *= np.exp(+1j * k[ax] * L[ax])
        assert np.linalg.norm(want - fb) < 1e-12 * np.linalg.norm(fb)


def test_periodic_sides_drop_their_pml_except_sigma_dz_quirk():
    """eff_boundaries zeroes the PML of Periodic / PEC / PMC sides; sigma_Dz keeps the raw
    thickness (Boundaries.jl:154-161)."""
    sim = kb.Simulation([2, 2, 2], [0, 0, 0], 10, [], boundaries=[[0.5, 0.5]] * 3,
                        boundary_conditions=[[kb.Periodic(), kb.Periodic()], [kb.PECBoundary(), kb.PML()], [kb.Periodic(), kb.Periodic()]])
    sim.host_prepare()
    sb, sd = sim.sigma
    assert not sb[0].any() and not sd[0].any()
    assert not sb[1][:10].any() and sb[1][-5:].any()
    assert not sb[2].any() and sd[2].any()


@pytest.mark.gpu
@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_gpu_periodic_xy_pml_z(dtype):
    bc = [[kb.Periodic(), kb.Periodic()], [kb.Bloch(0.0), kb.Bloch(0.0)], [kb.PML(), kb.PML()]]
    rng = np.random.default_rng(3)
    N = (36, 28, 44)
    eps = [(1.0 / rng.uniform(1.0, 3.0, N)).astype(dtype) for _ in range(3)]
    p = Pair([3.6, 2.8, 4.4], 10, [0.0, 0.0, 1.0], dtype, eps_inv=eps, boundary_conditions=bc,
             sources=[(kb.EX, [0, 0, -0.8], [3.6, 2.8, 0], CW), (kb.HY, [1.2, -1.0, 0.3], [0, 0, 0], CW)],
             monitors=[(kb.EX, [0, 0, 0.9], [3.6, 2.8, 0], [1.0], 1), (kb.HZ, [0, 0, 0], [0, 2.8, 4.4], [0.9, 1.1], 1)])
    p.step(80)
    tol = 1e-5 if dtype is np.float32 else 1e-12
    assert p.total_field_error() < tol
    for km, om in zip(p.kmon, p.omon):
        assert rel_l2(p.k.get_dft(km), p.o.get_dft(om)) < tol


@pytest.mark.gpu
def test_gpu_all_periodic_translation_and_oracle():
    def run(center):
        p = Pair([2.0, 1.5, 1.0], 8, [0.0, 0.0, 0.0], np.float32, boundary_conditions=PER,
                 sources=[(kb.EZ, center, [0, 0, 0], CW), (kb.HX, center, [0, 0, 0], CW)])
        p.step(40)
        return p
    a = run([0.0, 0.0, 0.0])
    assert a.total_field_error() < 1e-5
    b = run([0.5, -0.375, 0.25])
    for c in range(6):
        assert np.array_equal(np.roll(a.k.get_field(c), (4, -3, 2), axis=(0, 1, 2)), b.k.get_field(c))


@pytest.mark.gpu
def test_gpu_periodic_z_with_sigma_dz_quirk():
    """z periodic while a PML thickness is still given on z: sigma_Bz is dropped, sigma_Dz is not.
    (E damped, H not, on a torus: the configuration grows exponentially — 1e5 after 70 steps in
    both the oracle and the CUDA path — so the comparison stops after 25 steps.)"""
    bc = [[kb.PML(), kb.PML()], [kb.PML(), kb.PML()], [kb.Periodic(), kb.Periodic()]]
    p = Pair([3.0, 3.0, 2.4], 10, [0.8, 0.8, 0.6], np.float32, boundary_conditions=bc,
             sources=[(kb.EY, [0.1, 0, 0.2], [0, 0, 0], CW)], monitors=[(kb.EY, [0, 0, 0], [3.0, 0, 2.4], [1.0], 1)])
    p.step(25)
    assert p.total_field_error() < 1e-5
    assert rel_l2(p.k.get_dft(p.kmon[0]), p.o.get_dft(p.omon[0])) < 1e-5


def _bloch_pair(dtype, bc, pml, poles=()):
    return Pair([2.0, 1.6, 2.4], 10, pml, dtype, boundary_conditions=bc, poles=poles,
                sources=[(kb.EZ, [0.1, -0.2, 0.0], [0, 0, 0], CW), (kb.HY, [-0.3, 0.1, 0.2], [0.4, 0, 0], CW)],
                monitors=[(kb.EZ, [0, 0, 0.3], [2.0, 1.6, 0], [0.9, 1.1], 1), (kb.HX, [0, 0.2, 0], [2.0, 0, 1.0], [1.0], 1)])


def _check_complex(p, nsteps, tol):
    p.step(nsteps)
    num = den = 0.0
    for c in range(6):
        a = p.k.get_field(c, part="complex").astype(np.complex128)
        b = p.o.get_field(c) + 1j * p.o.get_field(c, which="imag")
        num += np.sum(np.abs(a - b) ** 2)
        den += np.sum(np.abs(b) ** 2)
    assert den > 0 and (num / den) ** 0.5 < tol, (num / den) ** 0.5
    for km, om in zip(p.kmon, p.omon):
        assert rel_l2(p.k.get_dft(km), p.o.get_dft(om)) < tol


@pytest.mark.gpu
@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_gpu_bloch_xy_pml_z(dtype):
    """Bloch(k) on x and y (complex fields as two real field sets, phase in the wrap), PML on z,
    complex DFT accumulation, a Drude block that the imaginary part must also carry."""
    bc = [[kb.Bloch(0.7), kb.Bloch(0.7)], [kb.Periodic(), kb.Bloch(-0.4)], [kb.PML(), kb.PML()]]
    sg = np.zeros((20, 16, 24), dtype=dtype)
    sg[5:12, 3:9, 10:15] = 1.5
    p = _bloch_pair(dtype, bc, [0.0, 0.0, 0.6], poles=[(0.0, 0.3, sg)])
    _check_complex(p, 90, 1e-5 if dtype is np.float32 else 1e-12)
    assert np.abs(p.k.get_field(kb.EZ, part="imag")).max() > 1e-3


@pytest.mark.gpu
def test_gpu_bloch_all_axes_and_translation_covariance():
    bc = [[kb.Bloch(0.7), kb.Bloch(0.7)], [kb.Bloch(-0.4), kb.Bloch(-0.4)], [kb.Bloch(1.1), kb.Bloch(1.1)]]
    p = _bloch_pair(np.float64, bc, [0.0, 0.0, 0.0])
    _check_complex(p, 60, 1e-12)


@pytest.mark.gpu
def test_gpu_bloch_k0_is_bit_identical_to_periodic():
    bc0 = [[kb.Bloch(0.0), kb.Bloch(0.0)], [kb.Periodic(), kb.Periodic()], [kb.PML(), kb.PML()]]
    bcp = [[kb.Periodic(), kb.Periodic()], [kb.Periodic(), kb.Periodic()], [kb.PML(), kb.PML()]]
    a = _bloch_pair(np.float32, bc0, [0.0, 0.0, 0.6])
    b = _bloch_pair(np.float32, bcp, [0.0, 0.0, 0.6])
    a.k.step(50)
    b.k.step(50)
    for c in range(6):
        assert np.array_equal(a.k.get_field(c), b.k.get_field(c))
        assert not a.k.get_field(c, part="imag").any()
    for ka, kb_ in zip(a.kmon, b.kmon):
        assert np.array_equal(a.k.get_dft(ka), b.k.get_dft(kb_))


@pytest.mark.gpu
def test_gpu_complex_field_restrictions():
    bc = [[kb.Bloch(0.7), kb.Bloch(0.7)], [kb.PML(), kb.PML()], [kb.PML(), kb.PML()]]
    chi3 = np.zeros((20, 16, 24), dtype=np.float32)
    chi3[8:12, 6:10, 10:14] = 1.0
    with pytest.raises(kb.KhronosError, match="chi3"):
        Pair([2.0, 1.6, 2.4], 10, [0.0, 0.4, 0.4], np.float32, boundary_conditions=bc, chi3=chi3,
             sources=[(kb.EZ, [0, 0, 0], [0, 0, 0], CW)])


@pytest.mark.parametrize("k", [2.0, -1.2])
def test_oracle_bloch_band_frequencies_of_vacuum(k):
    """Physics anchor for Bloch(k): a vacuum cell of period L = 1 with f(x + L) = f(x) exp(i k L) supports
    plane waves with k_x = k + 2 pi m, i.e. resonances at f = |k + 2 pi m| / (2 pi).  A pulsed sheet source
    rings them up; the two lowest must appear in the spectrum of a point DFT monitor.  Pins the phase as
    exp(i k L) with k in radians per unit length (Chunking.jl:1745-1747) and the complex-field update."""
    import oracle as ko
    from bridge import oracle_from_simulation
    nthreads = ko.num_threads()
    ko.set_num_threads(2)            # 40 x 4 x 4 cells
    try:
        bc = [[kb.Bloch(k), kb.Bloch(k)], [kb.Periodic(), kb.Periodic()], [kb.Periodic(), kb.Periodic()]]
        freqs = np.linspace(0.05, 1.2, 461)
        src = kb.UniformSource(kb.GaussianPulseSource(fcen=0.5, fwidth=1.5), kb.EY, [0.13, 0, 0], [0, 5, 5])
        mon = kb.DFTMonitor(kb.EY, [-0.21, 0, 0], [0, 0, 0], list(freqs), decimation=2)
        sim = kb.Simulation([1.0, 0.1, 0.1], [0, 0, 0], 40, [src], boundaries=[[0, 0]] * 3, boundary_conditions=bc,
                            monitors=[mon], dtype=np.float64)
        o, m = oracle_from_simulation(sim)
        o.step(24000)                # t = 300: line width ~ 1 / 300
        s = np.abs(o.get_dft(m[0]).reshape(-1, len(freqs))).sum(0)
    finally:
        ko.set_num_threads(nthreads)
    for f_exp in (abs(k) / (2 * np.pi), (2 * np.pi - abs(k)) / (2 * np.pi)):
        win = np.abs(freqs - f_exp) < 0.03
        f_peak = freqs[win][np.argmax(s[win])]
        assert abs(f_peak - f_exp) < 0.004, (k, f_exp, f_peak)
        assert s[win].max() > 8 * np.median(s), (k, f_exp)
