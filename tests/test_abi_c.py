"""The C ABI as a Julia `ccall` binder would use it: a plain C++ program (tests/abi_c/abi_smoke.cpp)
that includes only include/khronos_b200.h and links libkhronos_b200.so — no Python, no ctypes in the
process.  CPU part: it compiles and links against the header/.so and fails loudly without a GPU.
GPU part: it registers a dipole + PML + per-voxel eps problem, steps it one khr_step per call, reads
fields and a DFT accumulator back and compares with vectors the oracle dumped."""
import os
import struct
import subprocess

import numpy as np
import pytest

import khronos_b200 as kb
from khronos_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "abi_c", "abi_smoke.cpp")


def _build(tmp_path):
    exe = str(tmp_path / "abi_smoke")
    libdir = os.path.dirname(_lib.LIB_PATH)
    gxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    subprocess.check_call([gxx, "-std=c++17", "-O1", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), SRC, "-o", exe,
                           "-L", libdir, "-lkhronos_b200", "-Wl,-rpath," + libdir, "-Wl,-rpath,/usr/local/cuda/lib64"])
    return exe


def _dump_case(path, nsteps=40):
    from bridge import oracle_from_simulation
    rng = np.random.default_rng(11)
    N = (36, 32, 28)
    e1 = (1.0 / rng.uniform(1.0, 3.0, N)).astype(np.float32)
    mon = kb.DFTMonitor(kb.EZ, [0, 0, 0.1], [2.0, 1.5, 0], [0.9, 1.1], 2)
    sim = kb.Simulation([3.6, 3.2, 2.8], [0, 0, 0], 10, [kb.UniformSource(kb.ContinuousWaveSource(1.0), kb.EZ, [0.1, 0, 0], [0, 0, 0])],
                        boundaries=[[0.8, 0.8]] * 3, monitors=[mon], eps_inv=[e1, e1, e1])
    o, mids = oracle_from_simulation(sim)       # also checks the host plan below against the oracle's own derivation
    sim.host_prepare()
    o.step(nsteps)
    sd = sim.source_data[0]
    with open(path, "wb") as f:
        f.write(struct.pack("8i", N[0], N[1], N[2], nsteps, sd["comp"], mon.component, len(mon.frequencies), mon.decimation))
        f.write(struct.pack("4d", *[float(v) for v in sim.grid.dl], float(sim.grid.dt)))
        for grp in range(2):
            for a in range(3):
                f.write(np.ascontiguousarray(sim.sigma[grp][a], dtype=np.float32).tobytes())
        f.write(struct.pack("6i", *sd["start"], *sd["dims"]))
        f.write(struct.pack("4d", *sd["src"].time_profile.params(np.float32)))
        amp = np.asarray(sd["amp"], dtype=np.complex64)
        f.write(np.ascontiguousarray(amp.transpose(2, 1, 0)).tobytes())          # x fastest, (re, im) interleaved
        f.write(struct.pack("6i", *mon.start, *mon.end))
        f.write(struct.pack("%dd" % len(mon.frequencies), *[float(np.float32(v)) for v in mon.frequencies]))
        f.write(np.ascontiguousarray(e1.transpose(2, 1, 0)).tobytes())
        for comp in (kb.EZ, kb.HX):
            f.write(np.ascontiguousarray(o.get_field(comp).astype(np.float32).transpose(2, 1, 0)).tobytes())
        f.write(np.ascontiguousarray(o.get_dft(mids[0]).astype(np.complex64).transpose(3, 2, 1, 0)).tobytes())


def test_c_client_compiles_links_and_fails_loudly_without_gpu(tmp_path):
    exe = _build(tmp_path)
    case = str(tmp_path / "case.bin")
    _dump_case(case, nsteps=4)
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        pytest.skip("GPU box: the GPU test below runs the client for real")
    out = subprocess.run([exe, case], capture_output=True, text=True, timeout=120)
    assert out.returncode == 2 and "no CUDA device" in out.stderr and "no CPU fallback" in out.stderr, out.stderr


@pytest.mark.gpu
def test_c_client_drives_the_abi_and_matches_the_oracle(tmp_path):
    exe = _build(tmp_path)
    case = str(tmp_path / "case.bin")
    _dump_case(case)
    out = subprocess.run([exe, case], capture_output=True, text=True, timeout=300)
    print(out.stdout)
    assert out.returncode == 0 and out.stdout.startswith("PASS"), out.stdout + out.stderr
