"""Grid, Yee-lattice index maps and PML profiles (host side, integer-exact).

Mirrors the Julia host code that stays in front of the time-step path; every
function cites the reference lines it follows (paths relative to
/root/reference).  The maps must be bit-exact, so the Julia type-promotion rules
(T∘T → T, T∘Float64 → Float64, Int∘T → T) are spelled out with explicit numpy
scalar types rather than left to numpy's own promotion.
"""
import math

import numpy as np

EX, EY, EZ, HX, HY, HZ, CENTER = range(7)
COMPONENT_NAMES = ("Ex", "Ey", "Ez", "Hx", "Hy", "Hz", "Center")

# src/utils.jl:26-38 get_component_voxel_count: +1 cell on the staggered axes
_STAGGER = (
    (0, 1, 1),  # Ex
    (1, 0, 1),  # Ey
    (1, 1, 0),  # Ez
    (1, 0, 0),  # Hx
    (0, 1, 0),  # Hy
    (0, 0, 1),  # Hz
    (0, 0, 0),  # Center
)


def component_stagger(comp):
    return _STAGGER[comp]


class Grid:
    """Derived grid quantities of SimulationData (src/DataStructures.jl:732-741).

    N = floor(L*res) in Float64; Δ = T(L/N); Δt = T(min(Δ)·Courant) with the
    product formed in Float64 (Courant is still the user's Float64 there).
    """

    def __init__(self, cell_size, cell_center, resolution, courant=0.5, dtype=np.float32, spacing=None):
        self.T = np.dtype(dtype).type
        T = self.T
        self.cell_size_user = [float(v) for v in cell_size]
        self.cell_center = [float(v) for v in cell_center]
        self.resolution = float(resolution)
        self.courant = float(courant)
        self.N = [int(math.floor(L * self.resolution)) for L in self.cell_size_user]
        if any(n < 1 for n in self.N):
            raise ValueError("3-D grids only: every axis needs at least one cell (2-D, Nz == 0, is not supported)")
        self.cell_size = [T(L) for L in self.cell_size_user]
        self.dl = [T(L / n) for L, n in zip(self.cell_size_user, self.N)]
        # non-uniform grid (DataStructures.jl:737-739): Δ of an axis given as one spacing per cell.
        # dl[a] becomes the representative scalar the reference uses outside the kernels
        # (_scalar_spacing(Δ) = Δ[1], utils.jl:4-5); Δt = min over all spacings * Courant (:692,740)
        self.dlv = [None, None, None]
        for a, v in enumerate(spacing or [None, None, None]):
            if v is None:
                continue
            v = np.asarray(v, dtype=T)
            if v.shape != (self.N[a],) or not np.all(v > 0):
                raise ValueError("grid spacing of axis %d must be %d positive values" % (a, self.N[a]))
            self.dlv[a] = v
            self.dl[a] = T(v[0])
        mins = [float(self.dl[a]) if self.dlv[a] is None else float(self.dlv[a].min()) for a in range(3)]
        self.dt = T(min(mins) * self.courant)

    # src/utils.jl:26-38
    def component_voxel_count(self, comp):
        st = _STAGGER[comp]
        return [self.N[a] + st[a] for a in range(3)]

    # src/utils.jl:139-154 get_yee_shift: -Δ/2 (computed in T, widened to Float64) on staggered axes
    def yee_shift(self, comp):
        T = self.T
        st = _STAGGER[comp]
        return [float(-self.dl[a] / T(2)) if st[a] else 0.0 for a in range(3)]

    # src/utils.jl:156-170 get_component_origin
    def component_origin(self, comp):
        T = self.T
        sh = self.yee_shift(comp)
        out = []
        for a in range(3):
            v = self.cell_center[a] - float(self.cell_size[a] / T(2))
            v = v + float(self.dl[a] / T(2))
            out.append(v + sh[a])
        return out

    # src/utils.jl:176-195 get_grid_idx (no rounding)
    def grid_idx(self, point, comp):
        T = self.T
        sh = self.yee_shift(comp)
        out = []
        for a in range(3):
            vmin = self.cell_center[a] - float(self.cell_size[a] / T(2))
            vmax = self.cell_center[a] + float(self.cell_size[a] / T(2))
            c2c = float(self.dl[a] / T(2))
            mn = (vmin + c2c) + sh[a]
            mx = (vmax - c2c) - sh[a]
            p = min(float(point[a]), mx)
            p = max(p, mn)
            idx = (p - mn) / float(self.dl[a]) + 1
            if math.isnan(idx):
                idx = 0.0
            out.append(idx)
        return out

    # src/utils.jl:103-113 GridVolume(sim, volume, component); :201-211 lower/upper idx
    def grid_volume(self, center, size, comp):
        lo = [float(c) - float(s) / 2 for c, s in zip(center, size)]
        hi = [float(c) + float(s) / 2 for c, s in zip(center, size)]
        a = self.grid_idx(lo, comp)
        b = self.grid_idx(hi, comp)
        start = [int(math.floor(v)) for v in a]
        end = [int(math.ceil(v)) for v in b]
        return start, end

    # ------------------------------------------------------------------ PML
    # src/Boundaries.jl:23-38 sigma_helper
    def _sigma_helper(self, idx, Ns, dx, length_left, length_right):
        T = self.T
        dt = self.dt

        def u0(L):
            den = T(4) * L          # 4 * pml_length * 1 / 3, left to right in T
            den = den * T(1)
            den = den / T(3)
            return (-math.log(1e-15) / float(den)) * (0.5 * float(dt))

        def u(x):
            x2 = x * x              # T
            sgn = T(1) if x > 0 else (T(-1) if x < 0 else T(0))
            return (float(x2) * 0.5) * float(sgn + T(1))

        total = T(Ns) * dx / T(2)
        real_idx = T(idx) * dx / T(2)
        if real_idx < length_left:
            return T(u0(length_left) * u((length_left - real_idx) / length_left))
        if (total - real_idx) < length_right:
            return T(u0(length_right) * u((length_right - (total - real_idx)) / length_right))
        return T(0)

    # sigma_helper for a spacing vector (Boundaries.jl:23-38 with _pml_total_length / _pml_position,
    # :44-62): positions accumulate in Float64, the total length is a sum in T (left to right here;
    # Julia's sum may reassociate, so these profiles match the reference to rounding only)
    def _sigma_helper_nu(self, idx, Ns, dxv, length_left, length_right):
        T = self.T
        dt = self.dt

        def u0(L):
            den = T(4) * L
            den = den * T(1)
            den = den / T(3)
            return (-math.log(1e-15) / float(den)) * (0.5 * float(dt))

        def u(x):
            sgn = 1.0 if x > 0 else (-1.0 if x < 0 else 0.0)
            return ((x * x) * 0.5) * (sgn + 1.0)

        n = len(dxv)
        total = T(0)
        for k in range(min(Ns // 2, n)):
            total = T(total + dxv[k])
        cell, frac = idx // 2, (idx % 2) * 0.5
        pos = 0.0
        for k in range(1, min(cell, n) + 1):
            pos += float(dxv[k - 1])
        if cell < n:
            pos += frac * float(dxv[min(cell + 1, n) - 1])
        if pos < float(length_left):
            return T(u0(length_left) * u((float(length_left) - pos) / float(length_left)))
        if (float(total) - pos) < float(length_right):
            return T(u0(length_right) * u((float(length_right) - (float(total) - pos)) / float(length_right)))
        return T(0)

    # src/Boundaries.jl:64-72 compute_sigma: length 2N+1, zeros when both thicknesses are 0
    def compute_sigma(self, axis, length_left, length_right):
        T = self.T
        Ns = 2 * self.N[axis] + 1
        s = np.zeros(Ns, dtype=T)
        ll, lr = T(length_left), T(length_right)
        if ll != 0 or lr != 0:
            for idx in range(1, Ns + 1):
                if self.dlv[axis] is None:
                    s[idx - 1] = self._sigma_helper(idx, Ns, self.dl[axis], ll, lr)
                else:
                    s[idx - 1] = self._sigma_helper_nu(idx, Ns, self.dlv[axis], ll, lr)
        return s

    # src/Chunking.jl:642-643: PML cell counts per side (T-typed quotient, then ceil)
    def pml_cells(self, axis, length_left, length_right):
        T = self.T
        ll, lr = T(length_left), T(length_right)
        n = self.N[axis]
        left_end = int(math.ceil(float(ll / self.dl[axis]))) if ll > 0 else 0
        right_start = n - int(math.ceil(float(lr / self.dl[axis]))) + 1 if lr > 0 else n + 1
        return left_end, right_start


def _sq(x):  # Julia lowers x^2 to x*x (literal_pow); pow() may round differently
    return x * x


# src/utils.jl:486-541 _compute_interpolation_weight_fast (Float64 throughout)
def interpolation_weight(point, lo, hi, size, ndims, delta):
    weight = 1.0
    for dim in range(ndims):
        p, a, b, dl, sz = float(point[dim]), float(lo[dim]), float(hi[dim]), float(delta[dim]), float(size[dim])
        if p <= a - dl or p >= b + dl:
            return 0.0
        if sz == 0.0:
            weight *= 1 - min(abs(p - (a + b) * 0.5) / dl, 1.0)
        elif sz < dl:
            if a <= p <= b:
                weight *= 1 - 0.5 * _sq(1.0 - (p - a) / dl) - 0.5 * _sq(1.0 - (b - p) / dl)
            elif p <= a and abs(p - a) < dl:
                if b < p + dl:
                    weight *= 0.5 * _sq(1.0 - (a - p) / dl) - 0.5 * _sq(1.0 - (b - p) / dl)
                else:
                    weight *= 0.5 * _sq(1.0 - (a - p) / dl)
            elif p >= b and abs(p - b) < dl:
                if a > p - dl:
                    weight *= 0.5 * _sq(1.0 - (p - b) / dl) - 0.5 * _sq(1.0 - (p - a) / dl)
                else:
                    weight *= 0.5 * _sq(1.0 - (p - b) / dl)
        else:
            if p < a and abs(p - a) < dl:
                weight *= 0.5 * _sq(1.0 - (a - p) / dl)
            elif p >= a and abs(p - a) < dl:
                weight *= 1 - 0.5 * _sq(1.0 - (p - a) / dl)
            elif p <= b and abs(p - b) < dl:
                weight *= 1 - 0.5 * _sq(1.0 - (b - p) / dl)
            elif p > b and abs(p - b) < dl:
                weight *= 0.5 * _sq(1.0 - (p - b) / dl)
    return weight
