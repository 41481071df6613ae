"""GPU: every execution-mode switch of the library computes the same numbers.

The switches only change how the same arithmetic is scheduled (streams vs. programmatic
dependent launch with per-z-chunk dependency counters, axis-specialised PML kernels, skipping
the material loads on tiles with constant material, tile/z-segment sizes), so the results must
be bit-identical to the default mode, and within the north-star tolerance of the CPU oracle.
"""
import contextlib
import os

import numpy as np
import pytest

import khronos_b200 as kb
from common import Pair, rel_l2

pytestmark = pytest.mark.gpu

CW = kb.ContinuousWaveSource(fcen=1.0)


@contextlib.contextmanager
def env(**kw):
    old = {k: os.environ.get(k) for k in kw}
    os.environ.update({k: str(v) for k, v in kw.items()})
    try:
        yield
    finally:
        for k, v in old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v


def _scene(dtype=np.float32, nx=72):
    """Piecewise-constant eps with one random block (uniform and non-uniform tiles), PML on all
    axes, a Drude slab, a plane source and a point source, two DFT monitors."""
    rng = np.random.default_rng(7)
    N = (nx, 44, 48)
    eps = [np.full(N, 1.0, dtype=dtype) for _ in range(3)]
    for e in eps:
        e[:, :, 24:] = dtype(1.0 / 2.25)
        e[30:40, 10:20, 10:30] = (1.0 / rng.uniform(1.0, 4.0, (10, 10, 20))).astype(dtype)
    sg = np.zeros(N, dtype=dtype)
    sg[20:50, 12:30, 30:34] = 1.2
    kw = dict(sources=[(kb.EX, [0, 0, -0.9], [3.0, 2.0, 0], CW), (kb.HZ, [0.4, 0.2, 0.5], [0, 0, 0], CW)],
              monitors=[(kb.EX, [0, 0, 0.3], [5.0, 3.0, 0], [0.9, 1.0], 1), (kb.HY, [0, 0.1, 0], [4.0, 0, 3.0], [1.0], 2)],
              eps_inv=eps, poles=[(0.0, 0.3, sg)])
    return ([nx / 10.0, 4.4, 4.8], 10, [1.0, 1.0, 1.0], dtype), kw


def _run_gpu(nsteps, dtype=np.float32, nx=72, **envkw):
    args, kw = _scene(dtype, nx)
    with env(**envkw):
        p = Pair(*args, **kw)
    p.k.step(nsteps)
    p.k.sync()
    fields = [p.k.get_field(c).copy() for c in range(6)]
    dfts = [np.array(p.k.get_dft(m)) for m in p.kmon]
    p.k.close()
    return p, fields, dfts


MODES = [
    dict(KHR_CHAIN=1),
    dict(KHR_SWEEP=1),
    dict(KHR_SWEEP=1, KHR_ZSEG=3),
    dict(KHR_CHAIN=1, KHR_ZSEG=3),
    dict(KHR_AXIS_SPEC=1),
    dict(KHR_AXIS_SPEC=1, KHR_CHAIN=1),
    dict(KHR_UNIFORM_TILES=0),
    dict(KHR_ZIGZAG=1),
    dict(KHR_SPLIT_UNIFORM=0),
    dict(KHR_SPLIT_UNIFORM=0, KHR_AXIS_SPEC=1),
    dict(KHR_MULTI_STREAM=0),
    dict(KHR_ZSEG=5, KHR_ZSEG_FULL=1),
    dict(KHR_TAIL_ZN=2, KHR_SORT_ITEMS=1),
    # the TMA-staged persistent PML kernel (pml_tma.cuh): same cascade, different data movement
    dict(KHR_TMA=1),
    dict(KHR_TMA=1, KHR_SPLIT_UNIFORM=0),
    dict(KHR_TMA=1, KHR_UNIFORM_TILES=0),
    dict(KHR_TMA=1, KHR_TMA_STAGES=2),
    dict(KHR_TMA=1, KHR_ZSEG=3),
    dict(KHR_TMA=0),
    dict(KHR_TMA=1, KHR_FUSE=1),
    dict(KHR_TMA=1, KHR_FUSE=1, KHR_ZSEG=2),
    dict(KHR_TMA=1, KHR_FUSE=1, KHR_ZSEG=3, KHR_FUSE_LAG=2),
    dict(KHR_GRAPH=1),
    dict(KHR_GRAPH=1, KHR_TMA=1),
    dict(KHR_FULL_NOPML=1),
    dict(KHR_FULL_NOPML=1, KHR_FULL_SPLIT=1, KHR_ZSEG_FULL=2),
    dict(KHR_FIXUP=1),
    dict(KHR_FIXUP=1, KHR_TMA=1),
    dict(KHR_LOCAL_CUTS=0),
    dict(KHR_LOCAL_CUTS=0, KHR_TMA=1),
]


@pytest.fixture(scope="module")
def baseline():
    return _run_gpu(60)


@pytest.mark.parametrize("mode", MODES, ids=lambda m: ",".join("%s=%s" % kv for kv in m.items()))
def test_mode_is_bit_identical_to_default(baseline, mode):
    _, f0, d0 = baseline
    _, f1, d1 = _run_gpu(60, **mode)
    for c in range(6):
        assert np.array_equal(f0[c], f1[c]), ("field", c, rel_l2(f1[c], f0[c]))
    for a, b in zip(d0, d1):
        assert np.array_equal(a, b), ("dft", rel_l2(b, a))


@pytest.mark.parametrize("mode", [dict(KHR_TMA=1), dict(KHR_TMA=1, KHR_SPLIT_UNIFORM=0)], ids=lambda m: ",".join("%s=%s" % kv for kv in m.items()))
def test_tma_kernel_ragged_rows_bit_identical(mode):
    """Nx = 70 (not a multiple of 4): the last thread of a row owns 2 valid cells (RAGGED variants)."""
    _, f0, d0 = _run_gpu(40, nx=70, KHR_TMA=0)
    _, f1, d1 = _run_gpu(40, nx=70, **mode)
    for c in range(6):
        assert np.array_equal(f0[c], f1[c]), ("field", c, rel_l2(f1[c], f0[c]))
    for a, b in zip(d0, d1):
        assert np.array_equal(a, b), ("dft", rel_l2(b, a))


def test_step_graph_is_used_and_counts_kernels():
    """KHR_GRAPH=1: khr_step replays a CUDA graph after the first step (one host launch per step); stepping one call at a
    time, as run() does, gives the same bits as one call for all steps."""
    args, kw = _scene()
    with env(KHR_GRAPH=1):
        p = Pair(*args, **kw)
    p.k.step(30)
    p.k.sync()
    k, r = p.k.graph_info()
    assert k >= 4 and r == 29, (k, r)
    f_batch = [p.k.get_field(c).copy() for c in range(6)]
    p.k.reset_fields()
    for _ in range(30):
        p.k.step(1)
    p.k.sync()
    for c in range(6):
        assert np.array_equal(f_batch[c], p.k.get_field(c)), c
    p.k.close()


@pytest.mark.parametrize("bloch", [False, True])
def test_tma_kernel_with_periodic_and_bloch_axes(bloch):
    """The TMA half-step kernel next to the wrap kernels: x periodic (Bloch(k) when complex), y PML, z PML; the halo
    boxes then read ghost columns the wrap kernels filled.  Bit-identical to the LDG kernels."""
    N = (64, 40, 44)
    rng = np.random.default_rng(3)
    eps = [np.full(N, 1.0, dtype=np.float32) for _ in range(3)]
    for e in eps:
        e[20:44, 8:30, 14:30] = (1.0 / rng.uniform(1.0, 4.0, (24, 22, 16))).astype(np.float32)
    bx = [kb.Bloch(0.9), kb.Bloch(0.9)] if bloch else [kb.Periodic(), kb.Periodic()]
    kw = dict(sources=[(kb.EZ, [0.3, 0, 0.2], [0, 0, 0], CW), (kb.HY, [-2.9, 0.1, 0], [0, 1.0, 1.0], CW)],
              monitors=[(kb.EZ, [0, 0, 0.3], [6.4, 2.0, 0], [1.0], 1)], eps_inv=eps,
              boundary_conditions=[bx, [kb.PML(), kb.PML()], [kb.PML(), kb.PML()]])
    out = []
    for tma in (0, 1):
        with env(KHR_TMA=tma):
            p = Pair([6.4, 4.0, 4.4], 10, [[0.0, 0.0], [0.8, 0.8], [0.8, 0.8]], np.float32, **kw)
        p.k.step(50)
        p.k.sync()
        parts = ("real", "imag") if bloch else ("real",)
        out.append(([p.k.get_field(c, part).copy() for c in range(6) for part in parts], np.array(p.k.get_dft(p.kmon[0]))))
        if tma == 0:
            p.o.step(50)
            assert p.total_field_error() < 1e-5
        p.k.close()
    for a, b in zip(out[0][0], out[1][0]):
        assert np.array_equal(a, b)
    assert np.array_equal(out[0][1], out[1][1])
    assert any(np.abs(a).max() > 0 for a in out[0][0])


def test_default_mode_matches_oracle(baseline):
    p, f0, d0 = baseline
    p.o.step(60)
    num = den = 0.0
    for c in range(6):
        b = p.o.get_field(c).astype(np.float64)
        num += np.sum((f0[c].astype(np.float64) - b) ** 2)
        den += np.sum(b ** 2)
    assert np.sqrt(num / den) < 1e-5
    for a, om in zip(d0, p.omon):
        assert rel_l2(a, p.o.get_dft(om)) < 1e-5


def test_chain_mode_float64_matches_oracle():
    args, kw = _scene(np.float64)
    with env(KHR_CHAIN=1):
        p = Pair(*args, **kw)
    p.step(40)
    assert p.total_field_error() < 1e-12
    for km, om in zip(p.kmon, p.omon):
        assert rel_l2(p.k.get_dft(km), p.o.get_dft(om)) < 1e-12


def test_uniform_tiles_are_detected():
    """The planner must mark constant-material tiles (and only those): with a constant eps array
    every E tile is uniform; the kernel statistics expose the count."""
    N = (64, 32, 24)
    eps = [np.full(N, 0.5, dtype=np.float32) for _ in range(3)]
    p = Pair([6.4, 3.2, 2.4], 10, [0.5, 0.5, 0.5], np.float32, sources=[(kb.EZ, [0, 0, 0], [0, 0, 0], CW)], eps_inv=eps)
    st = [s for s in p.k.kernel_stats() if ",E," in s["name"]]
    assert st and all(s["uniform_ctas"] == s["ctas"] for s in st), st
    sth = [s for s in p.k.kernel_stats() if ",H," in s["name"]]
    assert all(s["uniform_ctas"] == 0 for s in sth)
    eps[1][10, 5, 7] = 0.25
    q = Pair([6.4, 3.2, 2.4], 10, [0.5, 0.5, 0.5], np.float32, sources=[(kb.EZ, [0, 0, 0], [0, 0, 0], CW)], eps_inv=eps)
    st = [s for s in q.k.kernel_stats() if ",E," in s["name"]]
    assert sum(s["ctas"] - s["uniform_ctas"] for s in st) == 1, st
    q.step(30)
    assert q.total_field_error() < 1e-5
