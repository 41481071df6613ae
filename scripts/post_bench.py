"""Throughput of the device post-processing kernels: near-to-far, mode overlap, flux.
   python scripts/post_bench.py [nobs_theta nobs_phi]"""
import os, sys, time, json
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path[:0] = [R]
import numpy as np
import khronos_b200 as kb

nth, nph = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (90, 36)
freqs = list(np.linspace(2.0, 2.4, 5))
fm = kb.Near2FarMonitor([0, 0, 0.6], [6.4, 6.4, 0], freqs, theta=np.linspace(0, np.pi / 2, nth), phi=np.linspace(0, 2 * np.pi, nph), r=1e4)
sim = kb.Simulation([7.0, 7.0, 2.5], [0, 0, 0], 40, [kb.UniformSource(kb.ContinuousWaveSource(2.2), kb.EY, [0, 0, -0.25], [0, 0, 0])],
                    boundaries=[[0.5, 0.5]] * 3, monitors=[fm], dtype=np.float32)
sim.prepare_simulation()
sim.step(200)
sim.sync()
n1 = min(m.end[0] - m.start[0] + 1 for m in fm.monitors)
n2 = min(m.end[1] - m.start[1] + 1 for m in fm.monitors)
def timed(f, reps=3):
    f(); best = 1e9
    for _ in range(reps):
        t0 = time.perf_counter(); out = f(); best = min(best, time.perf_counter() - t0)
    return best, out
t_n2f, EH = timed(lambda: sim.compute_far_field(fm))
nobs = nth * nph
greens = 4.0 * nobs * n1 * n2 * len(freqs)
t_flux, fl = timed(lambda: sim.get_flux(fm))
mode = np.random.default_rng(0).normal(size=(4, n1, n2, len(freqs))) + 0j
t_mo, _ = timed(lambda: sim.compute_mode_amplitudes(fm, mode))
t_host, _ = timed(lambda: [sim.get_dft(m) for m in fm.monitors], reps=1)
print(json.dumps(dict(surface=[n1, n2], nobs=nobs, nfreq=len(freqs), near2far_s=round(t_n2f, 4), green_evals=greens,
                      green_evals_per_s=greens / t_n2f, flux_s=round(t_flux, 5), mode_overlap_s=round(t_mo, 5),
                      readback_of_the_four_dft_arrays_s=round(t_host, 4), EH_norm=float(np.linalg.norm(EH)))))
