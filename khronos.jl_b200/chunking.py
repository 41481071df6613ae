"""Chunk plan and halo index maps (host side, integers only — bit-exact targets).

The B200 library keeps one contiguous array set per GPU, so the reference's
27-region PML grid is never *executed*; it is reproduced here because the plan,
its adjacency and its halo ranges are the index maps the north star requires to
match bit-exactly, and because the rank bands of the slab decomposition follow
the reference's own rounding rule.  Paths cite /root/reference/src.
"""
import math


def _jl_round(x):
    """Julia round(Int, x): ties to even (Python's round() has the same rule)."""
    return int(round(x))


def pml_grid_intervals(grid, boundaries, nranks=0):
    """Per-axis index intervals of Chunking.jl:622-716 (_pml_grid_regions).

    nranks > 0 reproduces the is_distributed() refinement of the longest interior
    interval (Chunking.jl:676-716).
    """
    intervals = []
    for axis in range(3):
        n = grid.N[axis]
        if boundaries is not None:
            left_end, right_start = grid.pml_cells(axis, boundaries[axis][0], boundaries[axis][1])
        else:
            left_end, right_start = 0, n + 1
        iv = []
        if left_end >= 1:
            iv.append((1, left_end))
        i_s, i_e = left_end + 1, right_start - 1
        if i_s <= i_e:
            iv.append((i_s, i_e))
        if right_start <= n:
            iv.append((right_start, n))
        if not iv:
            iv.append((1, n))
        intervals.append(iv)
    if nranks > 0:
        best_axis, best_len = -1, 0
        for axis in range(3):
            for s, e in intervals[axis]:
                if s > 1 and e < grid.N[axis]:
                    ln = e - s + 1
                    if ln > best_len:
                        best_len, best_axis = ln, axis
        if best_axis >= 0 and best_len > 0:
            new = []
            for s, e in intervals[best_axis]:
                if s > 1 and e < grid.N[best_axis] and (e - s + 1) == best_len:
                    total = e - s + 1
                    for k in range(1, nranks + 1):
                        sub_s = s + _jl_round((k - 1) * total / nranks)
                        sub_e = s + _jl_round(k * total / nranks) - 1
                        if sub_s <= sub_e:
                            new.append((sub_s, sub_e))
                else:
                    new.append((s, e))
            intervals[best_axis] = new
    return intervals


def pml_grid_regions(grid, boundaries, nranks=0):
    """Cartesian product of the intervals, x outer / z inner (Chunking.jl:718-737).

    Returns a list of (start[3], end[3]) with 1-based inclusive cell indices.
    """
    iv = pml_grid_intervals(grid, boundaries, nranks)
    out = []
    for sx, ex in iv[0]:
        for sy, ey in iv[1]:
            for sz, ez in iv[2]:
                out.append(([sx, sy, sz], [ex, ey, ez]))
    return out


def pml_overlaps_chunk_axis(grid, boundaries, region, axis):
    """Chunking.jl:108-138 _pml_overlaps_chunk_axis."""
    if boundaries is None:
        return False
    T = grid.T
    pl, pr = T(boundaries[axis][0]), T(boundaries[axis][1])
    if pl == 0 and pr == 0:
        return False
    left_end, right_start = grid.pml_cells(axis, pl, pr)
    s, e = region
    if left_end > 0 and s[axis] <= left_end:
        return True
    if right_start <= grid.N[axis] and e[axis] >= right_start:
        return True
    return False


def compute_adjacency(regions, ndims=3):
    """Chunking.jl:570-603 _compute_adjacency -> list of (i, j, axis), all 1-based."""
    adj = []
    n = len(regions)
    for i in range(n):
        for j in range(i + 1, n):
            si, ei = regions[i]
            sj, ej = regions[j]
            for axis in range(ndims):
                touches = (ei[axis] == sj[axis] - 1) or (ej[axis] == si[axis] - 1)
                if not touches:
                    continue
                ok = True
                for o in range(ndims):
                    if o == axis:
                        continue
                    if ei[o] < sj[o] or ej[o] < si[o]:
                        ok = False
                        break
                if ok:
                    adj.append((i + 1, j + 1, axis + 1))
    return adj


def overlap_halo_ranges(src, dst, axis, src_upper, dst_lower):
    """Chunking.jl:1863-1909 _make_overlap_halo_ranges.

    src/dst are (start[3], end[3]); returns two lists of three (first, last)
    chunk-local cell ranges (raw array index = cell index + 1 in Julia).
    """
    sdim = [src[1][d] - src[0][d] + 1 for d in range(3)]
    ddim = [dst[1][d] - dst[0][d] + 1 for d in range(3)]
    sr = [(1, sdim[d]) for d in range(3)]
    dr = [(1, ddim[d]) for d in range(3)]
    sr[axis] = (sdim[axis], sdim[axis]) if src_upper else (1, 1)
    dr[axis] = (0, 0) if dst_lower else (ddim[axis] + 1, ddim[axis] + 1)
    for d in range(3):
        if d == axis:
            continue
        lo = max(src[0][d], dst[0][d])
        hi = min(src[1][d], dst[1][d])
        if lo > hi:
            sr[d] = (1, 0)
            dr[d] = (1, 0)
        else:
            sr[d] = (lo - src[0][d] + 1, hi - src[0][d] + 1)
            dr[d] = (lo - dst[0][d] + 1, hi - dst[0][d] + 1)
    return sr, dr


def component_send_range(comp_dims, center_range):
    """Chunking.jl:2184-2196: clamp the upper end to the component extent."""
    return [(center_range[d][0], min(center_range[d][1], comp_dims[d])) for d in range(3)]


def component_recv_range(comp_dims, axis, center_range):
    """Chunking.jl:2198-2214: as send, but the split axis keeps the ghost index."""
    r = component_send_range(comp_dims, center_range)
    r[axis] = center_range[axis]
    return r


def aux_allocation_pattern(pml_flags, has_sigma_b=False, has_sigma_d=False):
    """Chunking.jl:1091-1122: which C/U/W arrays a chunk allocates.

    Returns a dict name -> [x, y, z] booleans for CB, UB, WB, CD, UD, WD:
    U <=> PML on next axis, W <=> PML on own axis, C <=> sigma and (next or prev).
    """
    p = [bool(v) for v in pml_flags]
    out = {k: [False] * 3 for k in ("CB", "UB", "WB", "CD", "UD", "WD")}
    for d in range(3):
        nx, pv = (d + 1) % 3, (d + 2) % 3
        out["UB"][d] = out["UD"][d] = p[nx]
        out["WB"][d] = out["WD"][d] = p[d]
        out["CB"][d] = has_sigma_b and (p[nx] or p[pv])
        out["CD"][d] = has_sigma_d and (p[nx] or p[pv])
    return out


def chunk_sigma_slice(sigma_global, start, n):
    """Chunking.jl:1333-1345: chunk-local sigma copy, local[2i-1] = global[2(i+start-1)-1]."""
    loc = [sigma_global.dtype.type(0)] * (2 * n + 1)
    for i in range(1, n + 1):
        gi = 2 * (i + start - 1) - 1
        if 1 <= gi <= len(sigma_global):
            loc[2 * i - 2] = sigma_global[gi - 1]
    return loc


# cost of one voxel per time step relative to an interior voxel, measured on B200 with this
# library's kernels (profiles/r02_slab_cost_calibration.txt): a voxel inside the PML costs PML_COST
# on its first PML axis and EXTRA_AXIS_COST more per further axis.  The reference's model is the
# same idea with one factor, chunk_cost: "PML ~60% more expensive" (Chunking.jl:269-277).
PML_COST = 2.2
EXTRA_AXIS_COST = 0.4


def plane_costs(grid, boundaries, pml_cost=None, extra_axis_cost=None):
    """Relative cost of every z plane (1-based plane k -> costs[k-1]) from the PML voxel census."""
    pml_cost = PML_COST if pml_cost is None else pml_cost
    extra = EXTRA_AXIS_COST if extra_axis_cost is None else extra_axis_cost
    n = grid.N
    frac = []          # PML cells per axis
    zpml = [False] * n[2]
    for a in range(3):
        if boundaries is not None:
            left_end, right_start = grid.pml_cells(a, boundaries[a][0], boundaries[a][1])
        else:
            left_end, right_start = 0, n[a] + 1
        cnt = min(n[a], max(left_end, 0) + max(n[a] - right_start + 1, 0))
        frac.append(cnt)
        if a == 2:
            for k in range(1, n[2] + 1):
                zpml[k - 1] = k <= left_end or k >= right_start
    px, py = frac[0], frac[1]
    nx, ny = n[0], n[1]
    c0 = (nx - px) * (ny - py)                 # cells of a plane with no x / y PML
    c1 = px * (ny - py) + (nx - px) * py       # one of x / y
    c2 = px * py                               # both
    plain = c0 * 1.0 + c1 * pml_cost + c2 * (pml_cost + extra)
    inz = c0 * pml_cost + c1 * (pml_cost + extra) + c2 * (pml_cost + 2 * extra)
    return [inz if z else plain for z in zpml]


def z_slab_partition(grid, boundaries, nranks, rule="cost", pml_cost=None, extra_axis_cost=None):
    """One z slab per rank.  Returns a list of (z_start, nz) per rank, 1-based.

    rule="cost" (default): the z axis is cut by cumulative per-plane cost so that every rank gets
    the same work — the reference's assign_chunks_to_ranks partitions its chunks "into nranks bands
    by cumulative cost" the same way (Distributed.jl:104-148, chunk_cost Chunking.jl:269-277); with
    z-PML planes ~1.5x as expensive as the others, the end ranks get fewer planes.

    rule="reference": the literal index rule of the distributed PML-grid planner — the interior z
    interval is cut with sub_s = s + round((k-1)·len/n), sub_e = s + round(k·len/n) − 1
    (Chunking.jl:703-706) and the z-PML cells stay with the first / last rank.
    """
    n = grid.N[2]
    if nranks <= 1:
        return [(1, n)]
    if n < nranks:
        raise ValueError("fewer z cells than ranks")
    if rule == "reference":
        if boundaries is not None:
            left_end, right_start = grid.pml_cells(2, boundaries[2][0], boundaries[2][1])
        else:
            left_end, right_start = 0, n + 1
        s, e = left_end + 1, right_start - 1
        if e - s + 1 < nranks:   # interior too thin to give every rank a part: cut the whole axis
            s, e = 1, n
        total = e - s + 1
        cuts = []
        for k in range(1, nranks + 1):
            sub_s = s + _jl_round((k - 1) * total / nranks)
            sub_e = s + _jl_round(k * total / nranks) - 1
            cuts.append([sub_s, sub_e])
        cuts[0][0] = 1
        cuts[-1][1] = n
        return [(a, b - a + 1) for a, b in cuts]
    if rule != "cost":
        raise ValueError("rule must be 'cost' or 'reference'")
    costs = plane_costs(grid, boundaries, pml_cost, extra_axis_cost)
    cum = [0.0]
    for c in costs:
        cum.append(cum[-1] + c)
    total = cum[-1]
    bounds = [0]     # planes 1..bounds[k] belong to ranks < k
    for k in range(1, nranks):
        target = total * k / nranks
        lo = bounds[-1] + 1                  # at least one plane per rank ...
        hi = n - (nranks - k)                # ... and enough planes left for the ranks above
        best = min(range(lo, hi + 1), key=lambda j: (abs(cum[j] - target), j))
        bounds.append(best)
    bounds.append(n)
    return [(bounds[k] + 1, bounds[k + 1] - bounds[k]) for k in range(nranks)]
