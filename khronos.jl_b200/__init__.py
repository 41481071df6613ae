"""khronos.jl_b200 — B200-native (sm_100a) FDTD time-step path of Khronos.jl.

The package holds only what the hot path needs: `csrc/` (the CUDA kernels and
the C ABI declared in include/khronos_b200.h), the ctypes binding, and the
host-side mirror of the reference's Simulation / run / monitor interface.
Import as `khronos_b200` (see khronos_b200.py at the repository root — the
directory name carries a dot and cannot be imported directly).
"""
from . import _lib, chunking, grid
from .grid import EX, EY, EZ, HX, HY, HZ, Grid, interpolation_weight
from .simulation import (Absorber, Ball, ContinuousWaveSource, Cuboid, Cylinder, CustomSource, DFTMonitor, DrudeSusceptibility,
                         FluxMonitor, GaussianPulseSource, LorentzianSusceptibility, Material, ModeMonitor, Near2FarMonitor, DiffractionMonitor,
                         Object, Simulation, UniformSource, run, run_benchmark, step, stop_when_dft_decayed, PML, Periodic, Bloch, PECBoundary, PMCBoundary)
from ._lib import KhronosError, build
from . import workloads

__all__ = [
    "Absorber", "Ball", "ContinuousWaveSource", "Cuboid", "Cylinder", "CustomSource", "DFTMonitor", "DrudeSusceptibility",
    "FluxMonitor", "GaussianPulseSource", "LorentzianSusceptibility", "Material", "Object", "Simulation",
    "UniformSource", "run", "run_benchmark", "step", "stop_when_dft_decayed", "Grid", "interpolation_weight", "KhronosError", "build",
    "EX", "EY", "EZ", "HX", "HY", "HZ", "chunking", "grid", "workloads", "PML", "Periodic", "Bloch", "PECBoundary",
    "PMCBoundary", "Near2FarMonitor", "ModeMonitor", "DiffractionMonitor",
]
