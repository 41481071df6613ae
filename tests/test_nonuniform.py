"""Non-uniform grids (DataStructures.jl:737-739 Δ as a vector, Helpers.jl:283-291 get_inv_dx,
Boundaries.jl:44-62 PML position mapping).  CPU: the host mirror against the oracle and exact
degenerate cases; GPU: the NU kernel variants against the oracle."""
import numpy as np
import pytest

import khronos_b200 as kb
from common import Pair, rel_l2

CW = kb.ContinuousWaveSource(fcen=1.0)


def _graded(n, d0, amp, seed):
    rng = np.random.default_rng(seed)
    i = np.arange(n)
    return d0 * (1.0 + amp * np.sin(2 * np.pi * i / n + rng.uniform(0, 6.28)) + 0.05 * rng.uniform(-1, 1, n))


def test_constant_spacing_vector_is_bit_identical_to_uniform():
    src = [(kb.EZ, [0.1, 0, 0], [0, 0, 0], CW), (kb.HY, [0, 0.2, 0.1], [0, 0, 0], CW)]
    a = Pair([2.4, 2.0, 1.6], 10, None, np.float32, sources=src, build_gpu=False)
    g = a.grid
    vec = [np.full(g.N[ax], g.dl[ax], dtype=np.float32) for ax in range(3)]
    b = Pair([2.4, 2.0, 1.6], 10, None, np.float32, sources=src, build_gpu=False, grid_spacing=vec)
    assert b.grid.dt == a.grid.dt
    a.o.step(40)
    b.o.step(40)
    for c in range(6):
        assert np.array_equal(a.o.get_field(c), b.o.get_field(c))


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_nonuniform_host_maps_match_oracle(dtype):
    """dt = min(all spacings) * Courant; sigma profiles of the host mirror equal the oracle's bit for
    bit; a constant vector reproduces the uniform profile's left ramp to rounding."""
    N = (24, 20, 16)
    vec = [_graded(N[0], 0.1, 0.3, 1).astype(dtype), None, _graded(N[2], 0.1, 0.2, 2).astype(dtype)]
    p = Pair([2.4, 2.0, 1.6], 10, [0.5, 0.4, 0.3], dtype, grid_spacing=vec, build_gpu=False,
             sources=[(kb.EZ, [0, 0, 0], [0, 0, 0], CW)])
    g = p.grid
    assert float(g.dt) == float(dtype(min(float(vec[0].min()), float(g.dl[1]), float(vec[2].min())) * 0.5))
    assert float(g.dl[0]) == float(vec[0][0]) and float(g.dl[2]) == float(vec[2][0])   # _scalar_spacing
    p.k.host_prepare()
    for grp in range(2):
        for ax in range(3):
            assert np.array_equal(p.k.sigma[grp][ax], p.o.sigma(ax, grp)), (grp, ax)
    assert p.k.sigma[0][0].any() and p.k.sigma[0][2].any()
    u = kb.Grid([2.4, 2.0, 1.6], [0, 0, 0], 10, 0.5, dtype)
    c = kb.Grid([2.4, 2.0, 1.6], [0, 0, 0], 10, 0.5, dtype, spacing=[np.full(24, u.dl[0], dtype=dtype), None, None])
    su, sc = u.compute_sigma(0, 0.5, 0.5), c.compute_sigma(0, 0.5, 0.5)
    # left side: same positions.  (Right side: the vector form measures from sum(Δ) = L while the
    # scalar form uses (2N+1) Δ / 2 = L + Δ/2, Boundaries.jl:41-48, so that ramp sits half a cell
    # further in; both are the reference's formulas.)
    assert np.allclose(su[:24], sc[:24], rtol=1e-5 if dtype is np.float32 else 1e-12, atol=1e-9)
    assert sc[-1] > 0 and sc[-12:].any()


@pytest.mark.gpu
@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_gpu_nonuniform_grid_matches_oracle(dtype):
    """Graded spacing on all three axes, PML, per-voxel eps, a Drude block, two sources, DFT planes."""
    N = (44, 36, 40)
    vec = [_graded(N[a], 0.1, 0.25, 10 + a).astype(dtype) for a in range(3)]
    rng = np.random.default_rng(5)
    eps = [(1.0 / rng.uniform(1.0, 3.0, N)).astype(dtype) for _ in range(3)]
    sg = np.zeros(N, dtype=dtype)
    sg[18:28, 14:22, 16:26] = 1.2
    p = Pair([4.4, 3.6, 4.0], 10, [1.0, 0.8, 1.0], dtype, grid_spacing=vec, eps_inv=eps, poles=[(0.0, 0.3, sg)],
             sources=[(kb.EZ, [0.1, 0, 0], [0, 0, 0], CW), (kb.HX, [-0.4, 0.2, 0.1], [0, 0.6, 0], CW)],
             monitors=[(kb.EZ, [0, 0, 0.2], [4.4, 3.6, 0], [0.9, 1.1], 1), (kb.HY, [0.3, 0, 0], [0, 3.6, 4.0], [1.0], 1)])
    p.step(80)
    tol = 1e-5 if dtype is np.float32 else 1e-12
    assert p.total_field_error() < tol, p.field_errors()
    for km, om in zip(p.kmon, p.omon):
        assert rel_l2(p.k.get_dft(km), p.o.get_dft(om)) < tol


@pytest.mark.gpu
def test_gpu_nonuniform_differs_from_uniform_and_one_axis_only():
    """Only z graded (x, y stay scalar): the library fills the uniform axes itself."""
    N = (40, 40, 40)
    vz = _graded(N[2], 0.1, 0.3, 3).astype(np.float32)
    src = [(kb.EZ, [0, 0, 0], [0, 0, 0], CW)]
    p = Pair([4, 4, 4], 10, [1.0, 1.0, 1.0], np.float32, grid_spacing=[None, None, vz], sources=src)
    q = Pair([4, 4, 4], 10, [1.0, 1.0, 1.0], np.float32, sources=src, build_gpu=False)
    p.step(60)
    assert p.total_field_error() < 1e-5
    assert p.grid.dt < q.grid.dt


def test_oracle_graded_mesh_reproduces_the_fresnel_slab():
    """Physics anchor for the non-uniform curl (inv(Δ[i]) of the updated cell for both half-steps,
    Helpers.jl:283-298): the slab transmission of tests/test_physics_anchor.py on a z mesh graded by
    +-30 % (same total length, eps volume-averaged over each node's dual cell from the physical node
    positions) matches the analytic curve as well as the uniform mesh does (0.031 both)."""
    import oracle as ko
    from bridge import oracle_from_simulation
    import test_physics_anchor as t
    nthreads = ko.num_threads()
    ko.set_num_threads(2)
    try:
        res, cell_xy, buffer, pml = 40, 0.1, 1.5, 1.0
        cell_z = t.THICK + 2 * buffer + 2 * pml
        nz = int(cell_z * res)
        dz = (1.0 / res) * (1 + 0.3 * np.sin(2 * np.pi * np.arange(nz) / nz * 3))
        dz *= cell_z / dz.sum()

        def flux(with_slab):
            fwidth = 2 * np.pi * 0.5 * (1 / 0.6 - 1 / 1.5)
            src = kb.UniformSource(kb.GaussianPulseSource(fcen=1.0, fwidth=fwidth), kb.EX, [0, 0, -t.THICK / 2 - buffer / 2],
                                   [cell_xy + 1, cell_xy + 1, 0])
            fm = kb.FluxMonitor([0, 0, t.THICK / 2 + buffer / 2], [cell_xy, cell_xy, 0], list(t.FREQS), decimation=2)
            eps_inv = None
            if with_slab:
                zk = -cell_z / 2 + np.concatenate([[0.0], np.cumsum(dz)[:-1]])          # Ex node positions
                lo, hi = zk - np.concatenate([[dz[0]], dz[:-1]]) / 2, zk + dz / 2          # dual cells
                f = np.clip((np.minimum(hi, t.THICK / 2) - np.maximum(lo, -t.THICK / 2)) / (hi - lo), 0, 1)
                e = (1 / (1 + (t.N_SLAB ** 2 - 1) * f))[None, None, :] * np.ones((4, 4, 1))
                eps_inv = [e, e, e]
            sim = kb.Simulation([cell_xy, cell_xy, cell_z], [0, 0, 0], res, [src], boundaries=[[0, 0], [0, 0], [pml, pml]],
                                boundary_conditions=[[kb.Periodic(), kb.Periodic()], [kb.Periodic(), kb.Periodic()],
                                                     [kb.PML(), kb.PML()]], monitors=[fm], dtype=np.float64,
                                grid_spacing=[None, None, dz], eps_inv=eps_inv)
            o, m = oracle_from_simulation(sim)
            o.step(int(60 / float(sim.grid.dt)))
            return o.flux(fm.normal, m)

        T = flux(True) / flux(False)
    finally:
        ko.set_num_threads(nthreads)
    assert np.max(np.abs(T - t.T_ANALYTIC)) < 0.045
