"""CPU tests (no GPU): index maps, PML profiles, plan — the product's host code and the
oracle against (a) the reference's own golden vectors (tests/golden/reference_tests.json)
and (b) each other, bit-exact."""
import json
import math
import os

import numpy as np
import pytest

import khronos_b200 as kb
from khronos_b200 import chunking
import oracle as ko

HERE = os.path.dirname(os.path.abspath(__file__))
G = json.load(open(os.path.join(HERE, "golden", "reference_tests.json")))
COMP = {"Ex": 0, "Ey": 1, "Ez": 2, "Hx": 3, "Hy": 4, "Hz": 5}
DT = {"f32": np.float32, "f64": np.float64}


def _f(v):
    return [float("inf") if x == "inf" else float(x) for x in v]


def _grid(sim):
    return kb.Grid(sim["cell_size"], sim["cell_center"], sim["resolution"], 0.5, DT[sim["dtype"]])


def _oracle(sim, boundaries=None, dtype=None):
    return ko.OracleSim(dtype or DT[sim["dtype"]], sim["cell_size"], sim["cell_center"], sim["resolution"], 0.5, boundaries)


# ---------------------------------------------------------------- golden vectors of the reference
def test_grid_volume_golden():
    g = _grid(G["grid_volume"]["sim"])
    o = _oracle(G["grid_volume"]["sim"])
    for c in G["grid_volume"]["cases"]:
        start, end = g.grid_volume([0, 0, 0], _f(c["size"]), COMP[c["comp"]])
        assert start == c["start"] and end == c["end"], c
        assert [e - s + 1 for s, e in zip(start, end)] == c["N"]
        os_, oe = o.grid_volume([0, 0, 0], _f(c["size"]), COMP[c["comp"]])
        assert list(os_) == c["start"] and list(oe) == c["end"]


def test_source_footprint_golden():
    S = G["source_footprint"]
    for comp in range(6):
        sim = kb.Simulation(S["sim"]["cell_size"], S["sim"]["cell_center"], S["sim"]["resolution"],
                            [kb.UniformSource(kb.ContinuousWaveSource(1.0), comp, S["center"], [0, 0, 0])],
                            boundaries=[[1.0, 1.0]] * 3, dtype=np.float64)
        sim.host_prepare()
        amp = sim.source_data[0]["amp"]
        assert np.count_nonzero(amp) == S["point_voxels"]
        for axis in range(3):
            size = [0.0, 0.0, 0.0]
            size[axis] = float("inf")
            sim = kb.Simulation(S["sim"]["cell_size"], S["sim"]["cell_center"], S["sim"]["resolution"],
                                [kb.UniformSource(kb.ContinuousWaveSource(1.0), comp, S["center"], size)],
                                boundaries=[[1.0, 1.0]] * 3, dtype=np.float64)
            sim.host_prepare()
            amp = sim.source_data[0]["amp"]
            n = sim.grid.component_voxel_count(comp)[axis]
            assert np.count_nonzero(amp) == S["line_voxels_factor"] * n, (comp, axis)


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_pml_grid_plan_golden(dtype):
    P = G["pml_grid"]
    sim = dict(P["sim"], dtype="f32" if dtype is np.float32 else "f64")
    g = _grid(sim)
    regions = chunking.pml_grid_regions(g, P["boundaries_all"])
    assert len(regions) == P["regions_all"]
    total = 0
    for s, e in regions:
        vol = (e[0] - s[0] + 1) * (e[1] - s[1] + 1) * (e[2] - s[2] + 1)
        assert vol > 0 and min(s) >= 1 and all(e[a] <= g.N[a] for a in range(3))
        total += vol
    assert total == g.N[0] * g.N[1] * g.N[2]
    assert len(chunking.pml_grid_regions(g, P["boundaries_x_only"])) == P["regions_x_only"]
    flags = [[chunking.pml_overlaps_chunk_axis(g, P["boundaries_all"], r, a) for a in range(3)] for r in regions]
    cnt = [sum(f) for f in flags]
    assert cnt.count(0) == P["class_counts"]["interior"] and cnt.count(1) == P["class_counts"]["face"]
    assert cnt.count(2) == P["class_counts"]["edge"] and cnt.count(3) == P["class_counts"]["corner"]
    adj = chunking.compute_adjacency(regions)
    assert len(adj) == P["adjacencies"]
    interior = cnt.index(0) + 1
    assert sum(1 for (i, j, _) in adj if interior in (i, j)) == P["interior_neighbours"]
    # the oracle's planner gives the same maps
    o = _oracle(sim, P["boundaries_all"], dtype)
    oreg, ofl = o.plan_pml_grid()
    assert [list(r[:3]) for r in oreg] == [s for s, _ in regions]
    assert [list(r[3:]) for r in oreg] == [e for _, e in regions]
    assert [[bool(x) for x in f] for f in ofl] == flags
    assert [tuple(a) for a in ko.adjacency([s + e for s, e in regions])] == adj


def test_aux_allocation_pattern_golden():
    A = G["aux_pattern"]
    pat = chunking.aux_allocation_pattern(A["interior"]["pml"])
    assert not any(any(v) for v in pat.values())
    pat = chunking.aux_allocation_pattern(A["face_x"]["pml"])
    names = {k + "xyz"[d] for k, v in pat.items() for d in range(3) if v[d]}
    assert set(A["face_x"]["allocated_includes"]) <= names
    assert not (set(A["face_x"]["absent"]) & names)
    # oracle chunk allocation agrees for all 27 chunks
    P = G["pml_grid"]
    o = _oracle(P["sim"], P["boundaries_all"])
    o.prepare("chunked")
    assert o.num_chunks() == 27
    for q in range(27):
        info = o.chunk_info(q)
        pat = chunking.aux_allocation_pattern(info["pml"])
        for k, v in pat.items():
            for d in range(3):
                assert info["aux"][k + "xyz"[d]] == v[d]


def test_chunk_sigma_zero_on_non_pml_axes_golden():
    P = G["pml_grid"]
    g = _grid(P["sim"])
    o = _oracle(P["sim"], P["boundaries_all"])
    o.prepare("chunked")
    sig = [g.compute_sigma(a, 1.0, 1.0) for a in range(3)]
    seen = False
    for q in range(27):
        info = o.chunk_info(q)
        if list(info["pml"]) == [True, False, False]:
            seen = True
            for axis in G["chunk_sigma"]["face_x_zero_axes"]:
                assert np.all(o.chunk_sigma(q, 0, axis) == 0) and np.all(o.chunk_sigma(q, 1, axis) == 0)
            mine = chunking.chunk_sigma_slice(sig[0], int(info["start"][0]), int(info["n"][0]))
            assert np.array_equal(np.array(mine, dtype=np.float64), o.chunk_sigma(q, 0, 0).astype(np.float64))
        if not any(info["pml"]):
            assert o.chunk_sigma(q, 0, 0) is None
    assert seen


def test_ade_coefficients_golden():
    A = G["ade"]
    c = ko.ade_coefficients(A["lorentz"]["omega_0"], A["lorentz"]["gamma"], A["dt"])
    for k in ("gamma1", "gamma1_inv", "omega0_dt_sq"):
        assert math.isclose(c[k], A["lorentz"][k], rel_tol=1e-14)
    assert math.isclose(c["sigma_omega0_dt_sq"], A["lorentz"]["omega0_dt_sq"], rel_tol=1e-14) and not c["is_drude"]
    d = ko.ade_coefficients(0.0, A["drude"]["gamma"], A["dt"])
    assert d["is_drude"] and d["omega0_dt_sq"] == 0.0
    assert math.isclose(d["drude_coeff"], A["drude"]["drude_coeff"], rel_tol=1e-14)
    s = kb.DrudeSusceptibility(0.5, 3.0)
    assert isinstance(s, kb.LorentzianSusceptibility) and (s.omega_0, s.gamma, s.sigma) == (0.0, 0.5, 3.0)


def test_interpolation_weights_sum_golden():
    """test_interpolation.jl:42-92: weights of a point sum to 1, of a line/area to its measure."""
    tol = G["interpolation"]["tolerance"]
    d = 0.1
    xs = np.arange(-2.0, 2.0 + 1e-9, d)

    def total(center, size):
        lo = [c - s / 2 for c, s in zip(center, size)]
        hi = [c + s / 2 for c, s in zip(center, size)]
        t = 0.0
        for x in xs:
            for y in xs:
                w = kb.interpolation_weight([x, y], lo, hi, size, 2, [d, d])
                assert w == ko.interp_weight([x, y, 0], lo + [0], hi + [0], list(size) + [0], 2, [d, d, d])
                t += w
        return t

    assert abs(total([0, 0], [0, 0]) - 1.0) < tol
    assert abs(total([-0.12, -0.26], [0, 0]) - 1.0) < tol
    assert abs(total([0.14, -0.21], [5 * d, 0]) * d - 5 * d) < tol
    assert abs(total([0.14, -0.21], [0, 5 * d]) * d - 5 * d) < tol
    for ix in np.arange(-1.6, 1.6 + 1e-9, 0.4):
        assert abs(total([ix * d, ix * d], [0.4, 0.5]) * d * d - 0.4 * 0.5) < tol
        assert abs(total([0.14, ix * d], [5.0 * d, 0.5 * d]) - 5.0 * 0.5) < 1e-11


def test_absorber_ramp_golden():
    """test_absorber.jl:22-82: ramp grows monotonically towards the boundary, zero in the interior."""
    ab = [[kb.Absorber(10, 3), kb.Absorber(10, 3)], None, None]
    sim = kb.Simulation([4, 3, 3], [0, 0, 0], 10, [], absorbers=ab, dtype=np.float64)
    sim.host_prepare()
    s = sim.material_arrays["sigma_D"][0]
    prof = s[:, 5, 5]
    assert np.all(np.diff(prof[:10]) < 0) and np.all(np.diff(prof[-10:]) > 0)
    assert np.all(prof[10:-10] == 0) and prof[0] > 0
    o = ko.OracleSim(np.float64, [4, 3, 3], [0, 0, 0], 10, 0.5, None)
    o.add_absorber(0, 0, 10, 3, 0.0)
    o.add_absorber(0, 1, 10, 3, 0.0)
    for kind in ("sigma_D", "sigma_B"):
        for d in range(3):
            assert np.array_equal(o.get_material_array(kind, d), sim.material_arrays[kind][d]), (kind, d)


# ---------------------------------------------------------------- product host code == oracle, bit-exact
@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_grid_and_sigma_bit_exact(dtype):
    rng = np.random.default_rng(7)
    for cell, res, pml in (([4.3, 3.7, 2.9], 10, [[0.5, 1.0], [0.0, 0.8], [0.7, 0.0]]),
                           ([8.0, 8.0, 8.0], 64, [[1.0, 1.0]] * 3), ([12.0, 6.0, 3.32], 40, [[1.0, 1.0]] * 3),
                           ([7.0, 7.0, 2.5], 40, [[0.5, 0.5]] * 3)):
        g = kb.Grid(cell, [0.1, -0.2, 0.05], res, 0.5, dtype)
        o = ko.OracleSim(dtype, cell, [0.1, -0.2, 0.05], res, 0.5, pml)
        assert tuple(g.N) == o.N
        assert [float(x) for x in g.dl] == list(o.dl) and float(g.dt) == o.dt
        for a in range(3):
            mine = g.compute_sigma(a, pml[a][0], pml[a][1])
            ref = o.sigma(a, 0)
            assert mine.dtype == dtype and np.array_equal(mine, ref), a
            assert np.array_equal(mine, o.sigma(a, 1))
        for comp in range(6):
            assert np.array_equal(np.array(g.component_origin(comp)), o.component_origin(comp))
            for _ in range(40):
                c = rng.uniform(-0.6, 0.6, 3) * np.array(cell)
                s = rng.uniform(0, 1.0, 3) * np.array(cell) * rng.integers(0, 2, 3)
                st, en = g.grid_volume(c, s, comp)
                ost, oen = o.grid_volume(c, s, comp)
                assert st == list(ost) and en == list(oen)
        for nr in (0, 2, 3, 8):
            mine = chunking.pml_grid_regions(g, pml, nr)
            oreg, _ = o.plan_pml_grid(nr)
            assert [s + e for s, e in mine] == [list(r) for r in oreg], nr


def test_halo_ranges_bit_exact():
    g = kb.Grid([10, 10, 10], [0, 0, 0], 10, 0.5, np.float32)
    pml = [[1.0, 1.0]] * 3
    regions = chunking.pml_grid_regions(g, pml, 2)
    adj = chunking.compute_adjacency(regions)
    assert len(adj) > 54
    for (i, j, axis) in adj:
        for (a, b, up) in ((i, j, True), (j, i, False)):
            src, dst = regions[a - 1], regions[b - 1]
            lower_first = src[1][axis - 1] == dst[0][axis - 1] - 1
            sr, dr = chunking.overlap_halo_ranges(src, dst, axis - 1, lower_first, lower_first)
            osr, odr = ko.halo_ranges(src[0] + src[1], dst[0] + dst[1], axis - 1, lower_first, lower_first)
            # the oracle returns raw 0-based first/last pairs: raw index == cell index
            assert [v for r in sr for v in r] == list(osr) and [v for r in dr for v in r] == list(odr)
    # component clamp (Chunking.jl:2184-2214)
    cr = [(1, 11), (3, 9), (10, 10)]
    assert chunking.component_send_range([10, 10, 10], cr) == [(1, 10), (3, 9), (10, 10)]
    assert chunking.component_recv_range([10, 10, 10], 2, [(1, 11), (3, 9), (11, 11)]) == [(1, 10), (3, 9), (11, 11)]


def test_z_slab_partition_rule():
    g = kb.Grid([4, 4, 16], [0, 0, 0], 10, 0.5, np.float32)
    pml = [[1.0, 1.0]] * 3
    for n in (1, 2, 3, 4, 8):
        slabs = chunking.z_slab_partition(g, pml, n, rule="reference")
        assert len(slabs) == n and slabs[0][0] == 1
        assert sum(nz for _, nz in slabs) == g.N[2]
        for (a, na), (b, _) in zip(slabs[:-1], slabs[1:]):
            assert a + na == b
        if n > 1:
            # interior cuts follow the reference's rounding (Chunking.jl:703-706)
            iv = chunking.pml_grid_intervals(g, pml, n)[2]
            assert [s for s, _ in iv[1:-1]][1:] == [z for z, _ in slabs[1:]]


def test_z_slab_partition_cost_balanced():
    """Default rule: cuts by cumulative per-plane cost (the reference's assign_chunks_to_ranks partitions by
    cumulative chunk_cost, Distributed.jl:104-148): every rank within 2 % of the mean work, contiguous
    cover, fewer planes on the ranks that hold the z-PML; degenerate cases stay valid."""
    for cell, res, pml, n in (([12, 6, 3.3 * 8], 40, [[1.0, 1.0]] * 3, 8), ([8, 8, 8 * 4], 16, [[1.0, 1.0]] * 3, 4),
                              ([4, 4, 16], 10, [[1.0, 1.0]] * 3, 3), ([4, 4, 16], 10, [[0.0, 0.0], [0.5, 0.5], [0.0, 2.0]], 2),
                              ([4, 4, 16], 10, None, 4)):
        g = kb.Grid(cell, [0, 0, 0], res, 0.5, np.float32)
        slabs = chunking.z_slab_partition(g, pml, n)
        assert len(slabs) == n and slabs[0][0] == 1 and sum(nz for _, nz in slabs) == g.N[2]
        for (a, na), (b, _) in zip(slabs[:-1], slabs[1:]):
            assert a + na == b and na >= 1
        costs = chunking.plane_costs(g, pml)
        work = [sum(costs[a - 1:a - 1 + nz]) for a, nz in slabs]
        tol = max(costs) / (sum(costs) / n)          # one plane of slack
        assert max(work) / (sum(work) / n) < 1 + tol, (slabs, work)
        if pml is not None and pml[2][0] > 0 and pml[2][1] > 0 and n > 2:
            assert slabs[0][1] < slabs[1][1] and slabs[-1][1] < slabs[-2][1]
            ref = chunking.z_slab_partition(g, pml, n, rule="reference")
            wref = [sum(costs[a - 1:a - 1 + nz]) for a, nz in ref]
            assert max(work) < max(wref)             # better balanced than the literal rule
    g = kb.Grid([1, 1, 0.4], [0, 0, 0], 10, 0.5, np.float32)
    assert chunking.z_slab_partition(g, None, 4) == [(1, 1), (2, 1), (3, 1), (4, 1)]
    with pytest.raises(ValueError):
        chunking.z_slab_partition(g, None, 5)


def test_time_source_matches_oracle():
    for dtype in (np.float32, np.float64):
        tp = kb.GaussianPulseSource(fcen=1.0, fwidth=0.4)
        assert tp.cutoff() > tp.peak_time > 0
        p = tp.params(dtype)
        a = ko.eval_time_source(dtype, 1, p, 3.7)
        assert abs(a) <= 1.0 and abs(a) > 0
        assert ko.eval_time_source(dtype, 1, p, tp.peak_time + tp.cutoff() + 1.0) == 0
        cw = ko.eval_time_source(dtype, 0, kb.ContinuousWaveSource(1.0).params(dtype), 0.25)
        assert abs(cw - (-1j)) < 1e-6


def test_auto_decimation_rule():
    """Monitors.jl:33-78: D = floor(1/(2 f_max dt)), applied to monitors left at 1."""
    sim = kb.Simulation([3, 3, 3], [0, 0, 0], 16, [kb.UniformSource(kb.ContinuousWaveSource(1.0), kb.EZ, [0, 0, 0], [0, 0, 0])],
                        monitors=[kb.DFTMonitor(kb.EZ, [0, 0, 0], [1, 1, 0], [1.0]), kb.DFTMonitor(kb.EX, [0, 0, 0], [1, 1, 0], [1.0], 3)])
    sim.host_prepare()
    assert [m.decimation for m in sim.dft_monitors] == [16, 3]


def test_stop_when_dft_decayed_predicate():
    """Simulation.jl:411-485 restated in the host mirror: per-monitor change relative to the largest
    change seen; monitors without signal are ignored; minimum / maximum runtime."""
    import khronos_b200 as kb

    class Stub:
        def __init__(self, series, dt=1.0):
            self.series, self.i, self.dt = series, -1, dt

        def advance(self):
            self.i += 1

        def round_time(self):
            return self.i * self.dt

        def monitor_norms(self):
            return [s[min(self.i, len(s) - 1)] for s in self.series]

    def run(stub, stop, nmax=100):
        for _ in range(nmax):
            stub.advance()
            if stop(stub):
                return stub.i
        return None

    # a monitor that rings up and settles, one that never sees signal
    a = [0.0, 1.0, 3.0, 4.0, 4.5, 4.75, 4.76, 4.76000001, 4.76000001]
    assert run(Stub([a, [0.0] * 9]), kb.stop_when_dft_decayed(tolerance=1e-6)) == 7     # |4.76000001 - 4.76| / 2 <= 1e-6
    assert run(Stub([a, [0.0] * 9]), kb.stop_when_dft_decayed(tolerance=1e-12)) == 8    # no change at all
    # literal quirk: checks start at minimum_runtime; a monitor that stopped changing before that never
    # shows a change, is never "active", and only maximum_runtime ends the run
    assert run(Stub([a]), kb.stop_when_dft_decayed(tolerance=1e-6, minimum_runtime=20.0)) is None
    assert run(Stub([a]), kb.stop_when_dft_decayed(tolerance=1e-6, minimum_runtime=20.0, maximum_runtime=30.0)) == 31
    assert run(Stub([a]), kb.stop_when_dft_decayed(tolerance=1e-6, minimum_runtime=3.0)) == 7
    assert run(Stub([[0.0] * 5]), kb.stop_when_dft_decayed(), nmax=30) is None          # never any signal
    assert run(Stub([[0.0, 1.0, 2.0, 4.0, 8.0, 16.0, 32.0, 64.0]]), kb.stop_when_dft_decayed(maximum_runtime=5.0), nmax=7) == 6
    # two monitors: the slower one decides
    b = [0.0, 0.0, 0.0, 1.0, 2.0, 2.5, 2.75, 2.875, 2.875, 2.875]
    assert run(Stub([a, b]), kb.stop_when_dft_decayed(tolerance=1e-6)) == 8
    with pytest.raises(ValueError):
        kb.stop_when_dft_decayed(minimum_runtime=2.0, maximum_runtime=1.0)


def test_reference_drude_chi1_stability_testset():
    """"Drude ADE eigenvalue stability (chi1 correction)" (test/test_dispersive.jl:74-160) replayed with the
    oracle's ADE coefficients (Susceptibility.jl:74-85) and the chi1 value the host mirror folds into
    eps^-1 (Geometry.jl:1236-1353): with the correction every eigenvalue of the update matrix stays on
    or inside the unit circle, without it the silver-like Drude medium of uled.jl is unstable."""
    import oracle as ko
    nm = 1e-3
    freq0 = 1.0 / (450 * nm)
    chi_target = (0.028 + 2.88j) ** 2 - 1.0
    w0 = 2 * np.pi * freq0
    Gamma_ang = -w0 * chi_target.imag / chi_target.real
    gamma_meep = Gamma_ang / (2 * np.pi)
    sigma_drude = -chi_target.real * (w0 ** 2 + Gamma_ang ** 2) / Gamma_ang
    dx = 1.0 / 40
    dt = 0.5 * dx / np.sqrt(3.0)
    Co = dt / dx
    c = ko.ade_coefficients(0.0, gamma_meep, dt)
    assert c["is_drude"]
    g1, g1i = c["gamma1"], c["gamma1_inv"]
    assert np.isclose(g1, 1 - gamma_meep * np.pi * dt) and np.isclose(g1i, 1 / (1 + gamma_meep * np.pi * dt))
    assert np.isclose(c["drude_coeff"], gamma_meep * 2 * np.pi * dt ** 2)
    d = c["drude_coeff"] * sigma_drude
    chi1 = g1i * d / 2                    # what simulation.py folds into eps_inv: g1i * (gamma 2 pi dt^2) / 2 * sigma

    def max_eig(eps_eff, S_list):
        worst = 0.0
        for S in S_list:
            M = np.array([[1 - (S ** 2 + g1i * d) / eps_eff, 1j * S / eps_eff, -g1 * g1i / eps_eff, g1 * g1i / eps_eff],
                          [1j * S, 1, 0, 0], [g1i * d, 0, 2 * g1i, -g1i * g1], [0, 0, 1, 0]], dtype=complex)
            worst = max(worst, np.abs(np.linalg.eigvals(M)).max())
        return worst

    S_1d = [2 * Co * np.sin(kf * np.pi / 2) for kf in np.linspace(0.01, 1.0, 200)]
    assert max_eig(1 + chi1, S_1d) <= 1.0 + 1e-12
    assert max_eig(1 + chi1, [np.sqrt(3.0)]) <= 1.0 + 1e-12      # 3-D worst case as the reference writes it
    assert max_eig(1.0, S_1d) > 1.0                               # unstable without the correction


def test_reference_susceptibility_formulas_are_the_anchor_formulas():
    """"Susceptibility evaluation" (test/test_dispersive.jl:40-58): the reference's chi(w) for Lorentz and
    Drude media — the closed forms tests/test_physics_anchor.py holds the time-stepper to."""
    import test_physics_anchor as t
    f = t.FREQS
    w = 2 * np.pi * f
    lor = 2.0 + 1.5 * (2 * np.pi * 1.2) ** 2 / ((2 * np.pi * 1.2) ** 2 - w ** 2 - 1j * w * (2 * np.pi * 0.2))
    dru = 1.5 + (-3.0 * (2 * np.pi * 0.5) / (w ** 2 + 1j * w * (2 * np.pi * 0.5)))
    assert np.allclose(lor, t.DISPERSIVE["lorentz"][1], rtol=1e-13)
    assert np.allclose(dru, t.DISPERSIVE["drude"][1], rtol=1e-13)


# ---------------------------------------------------------------- the oracle derives its own inputs
def _user_level_cases():
    from khronos_b200 import workloads as w
    cases = {
        "waveguide": w.waveguide_mode(res=10),
        "sphere": w.sphere(res=8, nfreq=3),
        "uled": w.uled(res=12),
        "metalens": w.metalens(nx=48, ny=48, nz=40, res=16, pml_cells=6, pillars=2, rotate=True),
        "dipole": w.dipole(40),
    }
    return cases


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("name", ["waveguide", "sphere", "uled", "metalens", "dipole"])
def test_oracle_derives_its_own_inputs(name, dtype):
    """oracle/bridge.py builds the oracle from the user-level description only (source volumes and
    amplitudes, geometry objects, susceptibilities, monitor volumes) and compares every derived input
    with the product's host_prepare() bit for bit: GridVolume boxes, interpolation weights x profile,
    raster, pole sigma + PML zeroing, chi1 fold, auto-decimation (Sources.jl:43-135, Geometry.jl:150-246,
    1059-1353, Monitors.jl:33-78)."""
    from khronos_b200 import workloads as w
    from bridge import oracle_from_simulation
    sim = w.build_simulation(_user_level_cases()[name], dtype)
    o, mids = oracle_from_simulation(sim, check=True)
    assert len(mids) == len(sim.dft_monitors)
    if name == "uled":
        assert o.num_poles() == 2 and np.count_nonzero(o.get_pole_sigma(0)) > 0
        # chi1 fold: eps_inv inside the metal differs from the raster value 1/eps = 1
        e = o.get_material_array("eps_inv", 0)
        assert np.any((o.get_pole_sigma(0) != 0) & (e != 1))


def test_oracle_input_check_detects_a_wrong_host_plan():
    """Mutation test of the check itself: a perturbed amplitude / decimation / pole sigma / absorber ramp in
    the product's plan must be reported."""
    from khronos_b200 import workloads as w
    from bridge import oracle_from_simulation
    for mutate in ("amp", "dec", "pole", "box"):
        sim = w.build_simulation(w.uled(res=12), np.float32)
        sim.host_prepare()
        if mutate == "amp":
            sim.source_data[0]["amp"] = sim.source_data[0]["amp"] * np.complex64(1.0000001)
        elif mutate == "dec":
            sim.dft_monitors[3].decimation += 1
        elif mutate == "pole":
            w0, g_, s = sim.poles[0]
            s = s.copy()
            s[np.nonzero(s)[0][0], np.nonzero(s)[1][0], np.nonzero(s)[2][0]] *= np.float32(1.000001)
            sim.poles[0] = (w0, g_, s)
        else:
            sim.dft_monitors[0].start = [sim.dft_monitors[0].start[0] + 1] + list(sim.dft_monitors[0].start[1:])
        with pytest.raises(AssertionError):
            oracle_from_simulation(sim, check=True)
        for m in sim.dft_monitors:   # monitors are shared objects of the description
            m.decimation = m.user_decimation


def test_pole_sigma_overlapping_objects_follow_reference():
    """Geometry.jl:1082-1125 _rasterize_pole_sigma!: an object without the pole is skipped, so a
    higher-priority dielectric cladding does NOT clear the sigma of a lower-priority metal where they
    overlap (eps_inv does take the cladding's value there); the first matching susceptibility wins."""
    from bridge import oracle_from_simulation
    drude = kb.DrudeSusceptibility(gamma=0.05, sigma=60.0)
    dup = kb.LorentzianSusceptibility(omega_0=0.0, gamma=0.05, sigma=7.0)      # same key: ignored (first match)
    geom = [kb.Object(kb.Cuboid([0, 0, 0.0], [1.0, 1.0, 1.0]), kb.Material(epsilon=2.0)),               # cladding first
            kb.Object(kb.Cuboid([0, 0, 0.3], [2.0, 2.0, 0.4]), kb.Material(epsilon=1.0, susceptibilities=[drude, dup]))]
    for dtype in (np.float32, np.float64):
        sim = kb.Simulation([4, 4, 3], [0, 0, 0], 10, [kb.UniformSource(kb.ContinuousWaveSource(1.0), kb.EZ, [0, 0, 0], [0, 0, 0])],
                            boundaries=[[0.5, 0.5]] * 3, geometry=geom, dtype=dtype)
        o, _ = oracle_from_simulation(sim, check=True)
        assert len(sim.poles) == 1
        s = sim.poles[0][2]
        xs, ys, zs = sim._coords(kb.EX)
        X, Y, Z = np.meshgrid(xs, ys, zs, indexing="ij")
        metal = (np.abs(X) <= 1.0) & (np.abs(Y) <= 1.0) & (np.abs(Z - 0.3) <= 0.2)
        clad = (np.abs(X) <= 0.5) & (np.abs(Y) <= 0.5) & (np.abs(Z) <= 0.5)
        assert np.all(s[metal] == dtype(60.0)) and np.all(s[~metal] == 0)
        assert np.any(metal & clad)                      # the overlap exists and keeps its sigma
        e = sim.material_arrays["eps_inv"][0]
        both = metal & clad
        # cladding eps (1/2) there, then the chi1 fold on top of it
        assert np.all(e[both] < dtype(0.5)) and np.all(e[clad & ~metal] == dtype(0.5))


def test_absorber_and_gaussian_source_inputs_derived_by_oracle():
    from bridge import oracle_from_simulation
    ab = [[kb.Absorber(8, 3), kb.Absorber(8, 2, 3.0)], None, [None, kb.Absorber(5, 3)]]
    prof = lambda pt, comp: np.exp(-(pt[1] / 0.4) ** 2) * (1 + 0.2 * pt[2]) + 0 * pt[0]
    srcs = [kb.UniformSource(kb.GaussianPulseSource(1.0, 0.35), kb.EY, [-0.3, 0, 0.05], [0, 1.5, 1.2], amplitude=0.7 - 0.2j, profile=prof),
            kb.UniformSource(kb.GaussianPulseSource(1.3, 0.5, start_time=0.3), kb.HZ, [0.2, 0.1, 0], [1.0, 0, float("inf")])]
    mons = [kb.FluxMonitor([0.5, 0, 0], [0, 1.0, 1.0], [0.9, 1.1]), kb.DFTMonitor(kb.HX, [0, 0, 0.2], [1, 1, 0], [1.0], 3)]
    for dtype in (np.float32, np.float64):
        sim = kb.Simulation([4, 3, 3], [0.1, 0, -0.05], 10, srcs, boundaries=[[0.0, 0.0], [0.6, 0.6], [0.5, 0.0]], absorbers=ab,
                            monitors=mons, dtype=dtype)
        o, mids = oracle_from_simulation(sim, check=True)
        assert sim.material_arrays["sigma_D"] is not None and len(mids) == 5
