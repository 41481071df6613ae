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


def test_bloch_nonzero_k_is_rejected():
    with pytest.raises(kb.KhronosError):
        kb.Simulation([2, 2, 2], [0, 0, 0], 10, [], boundaries=[[0, 0]] * 3,
                      boundary_conditions=[[kb.Bloch(k=0.3), kb.Bloch(k=0.3)], [kb.PML(), kb.PML()], [kb.PML(), kb.PML()]])


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
