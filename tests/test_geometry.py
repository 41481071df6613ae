"""Geometry rasterisation + subpixel smoothing (Geometry.jl:150-246, 450-605, 795-972).
CPU: the oracle's restatement against the host point sampler and against analytic fill fractions.
GPU: khr_geometry_rasterize against the oracle, then a run on the device-made arrays."""
import numpy as np
import pytest

import khronos_b200 as kb
from common import Pair, rel_l2
from bridge import oracle_from_simulation

CW = kb.ContinuousWaveSource(fcen=1.0)


def _scene():
    rot = [[np.cos(0.5), np.sin(0.5), 0.0], [-np.sin(0.5), np.cos(0.5), 0.0], [0.0, 0.0, 1.0]]
    return [kb.Object(kb.Ball([0.21, -0.13, 0.07], 0.83), kb.Material(epsilon=3.0)),
            kb.Object(kb.Cuboid([-0.4, 0.3, 0.1], [1.9, 0.7, 1.1], axes=rot), kb.Material(epsilon=2.2, mu=1.4, sigma_D=0.3)),
            kb.Object(kb.Cuboid([0.0, 0.0, -0.9], [3.0, 3.0, 0.5]), kb.Material(epsilon=5.76, sigma_B=0.1))]


def _scene_cyl():
    """Cylinders (GeometryPrimitives Cylinder(c, r, h, a), the shape of benchmark/periodic_bloch.jl:61 and of
    examples/uled.jl:166-216): z axis, tilted axis, one that pokes through a slab of lower priority."""
    t = np.array([0.3, -0.2, 0.9])
    return [kb.Object(kb.Cylinder([0.25, -0.15, 0.05], 0.62, 1.3, [0, 0, 1]), kb.Material(epsilon=1.0)),
            kb.Object(kb.Cylinder([-0.9, 0.6, -0.3], 0.45, 1.7, t), kb.Material(epsilon=4.0, sigma_D=0.2)),
            kb.Object(kb.Cuboid([0.0, 0.0, 0.0], [5.0, 5.0, 0.9]), kb.Material(epsilon=12.0))]


def _sim(dtype, rasterizer, smoothing=None, geometry=None, res=10):
    return kb.Simulation([4.0, 3.6, 3.2], [0, 0, 0], res, [kb.UniformSource(CW, kb.EZ, [1.2, 1.0, 0.9], [0, 0, 0])],
                         boundaries=[[0.6, 0.6]] * 3, geometry=geometry or _scene(), dtype=dtype, rasterizer=rasterizer,
                         subpixel_smoothing=smoothing)


def _oracle_arrays(sim):
    o, _ = oracle_from_simulation(sim)
    return o, {k: [o.get_material_array(k, d) for d in range(3)] for k in ("eps_inv", "mu_inv", "sigma_D", "sigma_B")}


def test_oracle_raster_equals_host_point_sampler():
    """Without smoothing the oracle's bounding-box painting (last object first, earlier ones win) and the
    host mirror's per-voxel `first object containing the point` give the same arrays (Float64: both
    form the coordinates in Float64)."""
    host = _sim(np.float64, "host")
    host.host_prepare()
    _, dev = _oracle_arrays(_sim(np.float64, "device"))
    for k in ("eps_inv", "mu_inv", "sigma_D", "sigma_B"):
        for d in range(3):
            assert np.array_equal(host.material_arrays[k][d], dev[k][d]), (k, d)
    assert len(np.unique(dev["eps_inv"][0])) == 4          # background + three materials: priorities exercised


def test_oracle_cylinder_raster_and_volume():
    """Cylinder: the oracle's painting equals the host point sampler bit for bit, and the painted / volume-averaged
    fill reproduces the analytic volume pi r^2 h of an axis-aligned and of a tilted cylinder."""
    for dtype in (np.float32, np.float64):
        host = _sim(dtype, "host", geometry=_scene_cyl())
        host.host_prepare()
        _, dev = _oracle_arrays(_sim(dtype, "device", geometry=_scene_cyl()))
        for k in ("eps_inv", "sigma_D"):
            for d in range(3):
                assert np.array_equal(host.material_arrays[k][d], dev[k][d]), (k, d, dtype)
    for axis in ([0, 0, 1], [0.3, -0.2, 0.9]):
        cyl = [kb.Object(kb.Cylinder([0.03, -0.02, 0.05], 0.9, 1.4, axis), kb.Material(epsilon=3.0))]
        mk = lambda mode: _oracle_arrays(kb.Simulation([3.2, 3.2, 3.2], [0, 0, 0], 10, [], geometry=cyl, dtype=np.float64,
                                                       rasterizer="device", subpixel_smoothing=mode))
        (o0, a0), (o1, a1) = mk(None), mk("volume")
        want = 1.0 + 2.0 * (np.pi * 0.9 ** 2 * 1.4) / (3.2 ** 3)
        assert min(o1.smoothed_voxels) > 800
        for d in range(3):
            stair = abs(np.sum(1.0 / a0["eps_inv"][d]) / 32 ** 3 - want)
            smooth = abs(np.sum(1.0 / a1["eps_inv"][d]) / 32 ** 3 - want)
            assert smooth < 1.5e-3 and stair < 3e-2, (axis, d, stair, smooth)   # e.g. z axis, Ez grid: staircase 1.4e-2 -> smoothed 3.6e-4
            ch = a1["eps_inv"][d] != a0["eps_inv"][d]
            assert np.all((a1["eps_inv"][d][ch] > 1 / 3.0 - 1e-12) & (a1["eps_inv"][d][ch] < 1.0 + 1e-12))


@pytest.mark.parametrize("mode", ["volume", "anisotropic"])
def test_oracle_smoothing_of_a_planar_interface_is_exact(mode):
    """A slab whose faces fall inside a voxel layer: the fill fraction of a planar interface is exact,
    so the smoothed eps^-1 must equal the closed forms — volume averaging 1/<eps>; anisotropic
    (Farjadpour 2006) <eps^-1> for the components parallel to the interface and 1/<eps> for the
    normal one (Geometry.jl:953-965)."""
    z0, z1, eps = -0.4321, 0.5678, 4.0
    slab = [kb.Object(kb.Cuboid([0, 0, (z0 + z1) / 2], [50.0, 50.0, z1 - z0]), kb.Material(epsilon=eps))]
    sim = kb.Simulation([1.6, 1.6, 3.2], [0, 0, 0], 10, [], geometry=slab, dtype=np.float64, rasterizer="device",
                        subpixel_smoothing=mode)
    o, a = _oracle_arrays(sim)
    for d, comp in enumerate((kb.EX, kb.EY, kb.EZ)):
        zs = sim.grid.component_origin(comp)[2] + np.arange(32) * 0.1
        f = np.clip((np.minimum(zs + 0.05, z1) - np.maximum(zs - 0.05, z0)) / 0.1, 0.0, 1.0)   # exact fill fraction
        eps_avg, eps_inv_harm = f * eps + (1 - f), f / eps + (1 - f)
        want = 1 / eps_avg if (mode == "volume" or d == 2) else eps_inv_harm
        got = a["eps_inv"][d][3, 5, :]
        # only layers whose centre-sampled value differs from a neighbour's are touched (:904-916)
        raw = np.where((zs >= z0) & (zs <= z1), 1 / eps, 1.0)
        iface = np.zeros(32, dtype=bool)
        iface[1:] |= raw[1:] != raw[:-1]
        iface[:-1] |= raw[:-1] != raw[1:]
        assert np.allclose(got[iface], want[iface], rtol=1e-12, atol=0), (d, got[iface], want[iface])
        assert np.array_equal(got[~iface], raw[~iface])
    assert sum(o.smoothed_voxels) == 3 * 4 * 16 * 16


def test_oracle_smoothing_of_a_sphere():
    """Interface voxels only, values between the two materials, mean permittivity within the
    planar-approximation bias (curvature: R = 10 cells) of the analytic sphere volume."""
    ball = [kb.Object(kb.Ball([0.03, -0.02, 0.05], 1.0), kb.Material(epsilon=3.0))]
    mk = lambda mode: _oracle_arrays(kb.Simulation([3.2, 3.2, 3.2], [0, 0, 0], 10, [], geometry=ball, dtype=np.float64,
                                                   rasterizer="device", subpixel_smoothing=mode))
    (o0, a0), (o1, a1), (o2, a2) = mk(None), mk("volume"), mk("anisotropic")
    assert o0.smoothed_voxels == [0, 0, 0] and min(o1.smoothed_voxels) > 1500 and o1.smoothed_voxels == o2.smoothed_voxels
    want = 1.0 + 2.0 * (4.0 / 3.0 * np.pi) / (3.2 ** 3)
    for d in range(3):
        assert abs(np.sum(1.0 / a1["eps_inv"][d]) / 32 ** 3 - want) < 1e-3
        for a in (a1, a2):
            ch = a["eps_inv"][d] != a0["eps_inv"][d]
            assert np.all((a["eps_inv"][d][ch] > 1 / 3.0 - 1e-12) & (a["eps_inv"][d][ch] < 1.0 + 1e-12))
        # anisotropic >= volume averaging: (1 - n^2) <eps^-1> + n^2 / <eps> >= 1 / <eps>
        assert np.all(a2["eps_inv"][d] >= a1["eps_inv"][d] - 1e-15)
    org = kb.Grid([3.2, 3.2, 3.2], [0, 0, 0], 10, 0.5, np.float64).component_origin(kb.EZ)
    X = np.stack(np.meshgrid(*[org[q] + np.arange(32) * 0.1 for q in range(3)], indexing="ij"))
    r = np.sqrt((X[0] - 0.03) ** 2 + (X[1] + 0.02) ** 2 + (X[2] - 0.05) ** 2)
    assert np.all(np.abs(r[a1["eps_inv"][2] != a0["eps_inv"][2]] - 1.0) < 0.15)


@pytest.mark.gpu
@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("mode", [None, "volume", "anisotropic"])
@pytest.mark.parametrize("scene", ["sphere+cuboids", "cylinders"])
def test_gpu_rasterizer_matches_oracle(dtype, mode, scene):
    sim = _sim(dtype, "device", mode, geometry=_scene() if scene == "sphere+cuboids" else _scene_cyl())
    o, want = _oracle_arrays(sim)
    sim.prepare_simulation()
    tol = 0 if mode is None else (2e-6 if dtype is np.float32 else 1e-12)
    for k in ("eps_inv", "mu_inv", "sigma_D", "sigma_B"):
        for d in range(3):
            if want[k][d] is None:
                continue                     # the scene does not need this kind (e.g. no mu / sigma_B in the cylinder scene)
            got = sim.get_material(k, d)
            if tol == 0 or k != "eps_inv":
                assert np.array_equal(got, want[k][d]), (k, d, np.argwhere(got != want[k][d])[:5])
            else:
                assert np.max(np.abs(got - want[k][d]) / np.abs(want[k][d])) <= tol, (k, d)
    assert sim.smoothed_voxels == o.smoothed_voxels
    if mode is not None:
        assert sum(sim.smoothed_voxels) > 1000


@pytest.mark.gpu
def test_gpu_run_on_device_rasterized_sphere():
    """benchmark/sphere.jl shape at 64^3 with the reference's default anisotropic smoothing, geometry made
    on the device, fields and DFT against the oracle."""
    ball = [kb.Object(kb.Ball([0, 0, 0], 1.25), kb.Material(epsilon=3.0))]
    p = Pair([6.4, 6.4, 6.4], 10, [1.0, 1.0, 1.0], np.float32, geometry=ball, rasterizer="device", subpixel_smoothing="anisotropic",
             sources=[(kb.EX, [0, 0, -2.0], [6.4, 6.4, 0], CW)], monitors=[(kb.EX, [0, 0, 1.8], [4, 4, 0], [1.0], 1)])
    p.step(120)
    assert p.total_field_error() < 1e-5, p.field_errors()
    assert rel_l2(p.k.get_dft(p.kmon[0]), p.o.get_dft(p.omon[0])) < 1e-5
    assert sum(p.k.smoothed_voxels) > 0


@pytest.mark.gpu
def test_gpu_rasterizer_argument_errors():
    with pytest.raises(kb.KhronosError, match="device"):
        kb.Simulation([4, 4, 4], [0, 0, 0], 10, [], geometry=_scene(), rasterizer="device",
                      absorbers=[[kb.Absorber(4), None], None, None]).host_prepare()


def test_reference_subpixel_testset_on_sphere():
    """"Subpixel smoothing on sphere" (test/test_subpixel.jl:38-153) replayed on the oracle's restatement:
    eps = 12 ball of radius 1 over an eps = 1 background cuboid, 4^3 cell at resolution 10."""
    geom = [kb.Object(kb.Ball([0, 0, 0], 1.0), kb.Material(epsilon=12.0)),
            kb.Object(kb.Cuboid([0, 0, 0], [100.0, 100.0, 100.0]), kb.Material(epsilon=1.0))]
    src = [kb.UniformSource(CW, kb.EZ, [0, 0, 0], [0, 0, 0])]

    def arrays(mode):
        sim = kb.Simulation([4.0, 4.0, 4.0], [0, 0, 0], 10, src, boundaries=[[1.0, 1.0]] * 3, geometry=geom,
                            dtype=np.float32, rasterizer="device", subpixel_smoothing=mode)
        return _oracle_arrays(sim)[1]["eps_inv"]

    inv_s, inv_b = np.float32(1.0 / 12.0), np.float32(1.0)
    near = lambda v, w: np.isclose(v, w, rtol=1e-4, atol=0)
    none = arrays(None)
    assert np.all(near(none[0], inv_s) | near(none[0], inv_b))                      # binary values
    for mode in ("volume", "anisotropic"):
        e = arrays(mode)
        inter = ~near(e[0], inv_s) & ~near(e[0], inv_b)
        assert inter.sum() > 0                                                      # intermediate values at interfaces
        c = e[0].shape[0] // 2 - 1
        assert near(e[0][c, c, c], inv_s) and near(e[0][0, 0, 0], inv_b)            # interior / exterior unchanged
        assert np.all((e[0] >= inv_s - 1e-5) & (e[0] <= inv_b + 1e-5))              # bounded
        if mode == "anisotropic":                                                   # components differ at the interface
            diff = inter & (~near(e[0], e[1]) | ~near(e[0], e[2]))
            assert diff.sum() > 0
