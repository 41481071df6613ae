"""GPU parity of the step! branches and post-processing added after the core path:
the Kerr chi3 correction (Dispersive.jl:127-173), the near-to-far transformation
(Near2Far.jl:40-371) and the mode-overlap sums (ModeMonitor.jl:345-515), all through the C ABI
against the CPU oracle on the same inputs."""
import numpy as np
import pytest

import khronos_b200 as kb
from common import Pair, rel_l2

pytestmark = pytest.mark.gpu

TOL = {np.float32: 1e-5, np.float64: 1e-12}
CW = kb.ContinuousWaveSource(fcen=1.0)


def _kerr_pair(dtype, with_pole=False, src_inside=False):
    N = (40, 40, 40)
    chi3 = np.zeros(N, dtype=dtype)
    chi3[12:30, 14:27, 13:29] = 0.8          # strong enough that 1/(1 + chi3 |E|^2) departs from 1 by ~1e-2
    chi3[20:24, 18:22, 18:22] = 2.5
    poles = []
    if with_pole:
        sg = np.zeros(N, dtype=dtype)
        sg[16:34, 10:24, 15:25] = 1.5        # overlaps the Kerr block only partly
        poles = [(0.0, 0.3, sg)]
    src_pos = [0.3, 0.1, 0.0] if src_inside else [-0.9, 0.0, 0.0]
    return Pair([4, 4, 4], 10, [1.0, 1.0, 1.0], dtype, chi3=chi3, poles=poles,
                sources=[(kb.EZ, src_pos, [0, 0, 0], CW), (kb.EX, [0.2, -0.1, 0.3], [0, 0, 0], CW)],
                monitors=[(kb.EZ, [0, 0, 0], [2, 2, 0], [1.0], 1)])


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("with_pole,src_inside", [(False, False), (True, False), (False, True), (True, True)])
def test_kerr_chi3_matches_oracle(dtype, with_pole, src_inside):
    p = _kerr_pair(dtype, with_pole, src_inside)
    p.step(90)
    assert p.total_field_error() < TOL[dtype], p.field_errors()
    a, b = p.k.get_dft(p.kmon[0]), p.o.get_dft(p.omon[0])
    assert rel_l2(a, b) < TOL[dtype]


def test_kerr_changes_the_fields():
    """The correction is not a no-op in this configuration (guards the tests above)."""
    p = _kerr_pair(np.float64)
    q = Pair([4, 4, 4], 10, [1.0, 1.0, 1.0], np.float64, build_gpu=False,
             sources=[(kb.EZ, [-0.9, 0, 0], [0, 0, 0], CW), (kb.EX, [0.2, -0.1, 0.3], [0, 0, 0], CW)])
    p.step(90)
    q.o.step(90)
    diff = rel_l2(p.k.get_field(kb.EZ), q.o.get_field(kb.EZ))
    assert diff > 1e-4, diff


def test_kerr_inside_pml_is_rejected():
    chi3 = np.zeros((40, 40, 40), dtype=np.float32)
    chi3[2:6, 10:20, 10:20] = 1.0   # x cells 3..6 lie in the 10-cell PML
    with pytest.raises(kb.KhronosError, match="PML"):
        Pair([4, 4, 4], 10, [1.0, 1.0, 1.0], np.float32, chi3=chi3, sources=[(kb.EZ, [0, 0, 0], [0, 0, 0], CW)])


def _plane_pair(dtype, normal, cls, **kw):
    size = [2.0, 2.0, 2.0]
    size[normal] = 0.0
    center = [0.0, 0.0, 0.0]
    center[normal] = 0.6
    fm = cls(center, size, [0.9, 1.0, 1.15], **kw)
    mons = [(m.component, m.center, m.size, m.frequencies, 1) for m in fm.monitors]
    p = Pair([4, 4, 4], 10, [1.0, 1.0, 1.0], dtype, sources=[(kb.EZ, [0, 0, 0], [0, 0, 0], CW),
                                                              (kb.EX, [0.2, -0.1, 0.3], [0, 0, 0], CW)], monitors=mons)
    p.step(60)
    fm.monitors = p.kmon
    return p, fm


@pytest.mark.parametrize("normal", [0, 1, 2])
@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_near2far_matches_oracle(normal, dtype):
    """khr_near2far against the oracle's restatement of _compute_far_field_cpu: near, intermediate
    and far zone points, both normal directions, a medium with eps, mu != 1."""
    rng = np.random.default_rng(7)
    obs = np.concatenate([rng.uniform(-3, 3, (5, 3)) + 4.0, rng.uniform(-1, 1, (4, 3)) * 50.0,
                          np.array([[1e4, 2e3, -5e3], [0.0, 0.0, 1e6]])])
    for ns, eps, mu in (("+", 1.0, 1.0), ("-", 2.25, 1.3)):
        p, fm = _plane_pair(dtype, normal, kb.Near2FarMonitor, normal_dir=ns, medium_eps=eps, medium_mu=mu)
        dev = p.k.compute_far_field(fm, obs)
        ora = p.o.near2far(normal, p.omon, fm.normal_sign, eps, mu, p.k._plane_bases(fm),
                           [float(f) for f in fm.frequencies], obs)
        assert dev.shape == ora.shape == (obs.shape[0], 6, 3)
        # per observation point: its six components jointly (single components may vanish by symmetry)
        for io in range(obs.shape[0]):
            assert rel_l2(dev[io], ora[io]) < TOL[dtype], (io, rel_l2(dev[io], ora[io]))


def test_near2far_theta_phi_grid_and_power():
    p, fm = _plane_pair(np.float64, 2, kb.Near2FarMonitor, theta=np.linspace(0, np.pi / 2, 7),
                        phi=np.linspace(0, 2 * np.pi, 5), r=1e3)
    EH = p.k.compute_far_field(fm)
    assert EH.shape == (35, 6, 3)
    ora = p.o.near2far(2, p.omon, 1.0, 1.0, 1.0, p.k._plane_bases(fm), fm.frequencies, p.k.far_field_points(fm))
    assert rel_l2(EH, ora) < 1e-10
    power = p.k.compute_far_field_power(EH, fm.theta, fm.phi)
    assert power.shape == (7, 5) and np.all(power >= 0) and power.max() > 0
    # far zone: E is transverse, |E_r| << |E|
    pts = p.k.far_field_points(fm)
    rhat = pts / np.linalg.norm(pts, axis=1, keepdims=True)
    er = np.abs(np.sum(EH[:, 0:3, 0] * rhat, axis=1))
    assert np.max(er) < 1e-2 * np.max(np.abs(EH[:, 0:3, 0]))


@pytest.mark.parametrize("normal", [0, 1, 2])
@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_mode_overlap_matches_oracle(normal, dtype):
    p, fm = _plane_pair(dtype, normal, kb.ModeMonitor)
    dft = [p.k.get_dft(m) for m in fm.monitors]
    t1, t2 = fm.tangential
    n1 = min(a.shape[t1] for a in dft)
    n2 = min(a.shape[t2] for a in dft)
    rng = np.random.default_rng(11)
    mode = rng.normal(size=(4, n1, n2, 3)) + 1j * rng.normal(size=(4, n1, n2, 3))
    ap, am = p.k.compute_mode_amplitudes(fm, mode)
    oap, oam, _ = p.o.mode_amplitudes(normal, p.omon, mode)
    assert rel_l2(ap, oap) < TOL[dtype] and rel_l2(am, oam) < TOL[dtype], (ap, oap, am, oam)
    # a mode equal to the recorded field itself has a+ = 1 exactly (overlap_plus = 4 P_mode)
    def avg(a):
        a = np.asarray(a, dtype=np.complex128)
        if a.shape[normal] >= 2:
            a = (np.take(a, 0, axis=normal) + np.take(a, 1, axis=normal)) / 2
        else:
            a = np.take(a, 0, axis=normal)
        return a[:n1, :n2, :]
    self_mode = np.stack([avg(a) for a in dft])
    ap1, _ = p.k.compute_mode_amplitudes(fm, self_mode)
    assert np.allclose(ap1.real, 1.0, atol=1e-5 if dtype is np.float32 else 1e-12), ap1
    with pytest.raises(kb.KhronosError):
        p.k.compute_mode_amplitudes(fm, mode[:, :-1])


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_diffraction_orders_match_oracle(dtype):
    """khr_diffraction (DiffractionMonitor.jl:87-165) on a binary grating with periodic x/y: same
    propagating set and the same order powers as the oracle's restatement."""
    from test_post_oracle import _grating
    p, fm = _grating(dtype, build_gpu=True)
    p.step(500)
    got = p.k.get_diffraction_efficiencies(fm, max_order=3)
    power, prop = p.o.diffraction(2, p.omon, 3, float(p.grid.cell_size[0]), float(p.grid.cell_size[1]),
                                  [float(f) for f in fm.frequencies])
    want = {(m - 3, n - 3): power[:, m, n] for m in range(7) for n in range(7) if prop[:, m, n].any()}
    assert set(got) == set(want) and len(got) == 5
    scale = max(np.abs(v).max() for v in want.values())
    tol = 1e-5 if dtype is np.float32 else 1e-12
    for k in want:
        assert np.max(np.abs(got[k] - want[k])) < tol * scale, (k, got[k], want[k])
