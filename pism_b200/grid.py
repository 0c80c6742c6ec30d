"""Grid description and PISM's 2D domain decomposition (reference: src/util/IceGrid.cc).

Only what the SIAFD path needs from IceGrid: spacing and coordinates (IceGrid.cc:588-659,
cell-corner registration, not periodic), vertical levels (:381-424), the processor-grid rule
(:443-484) and ownership ranges (:489-499).  Local arrays use the DMDA ghosted layout
[j][i][dof] (src/util/IceModelVec_inline.hh:28-40).
"""
import math
from dataclasses import dataclass

import numpy as np


def compute_vertical_levels(Lz, Mz, spacing="equal", lam=4.0):
    """IceGrid::compute_vertical_levels, IceGrid.cc:381-424."""
    if Mz < 2:
        raise ValueError("Mz must be at least 2")
    if Lz <= 0:
        raise ValueError("Lz must be positive")
    z = np.empty(Mz, dtype=np.float64)
    if spacing == "equal":
        dz = Lz / (float(Mz) - 1)
        for k in range(Mz - 1):
            z[k] = dz * float(k)
    elif spacing == "quadratic":
        if lam <= 0:
            raise ValueError("lambda must be positive")
        for k in range(Mz - 1):
            zeta = float(k) / (float(Mz) - 1)
            z[k] = Lz * ((zeta / lam) * (1.0 + (lam - 1.0) * zeta))
    else:
        raise ValueError("spacing can not be UNKNOWN")
    z[Mz - 1] = Lz
    return z


def compute_nprocs(Mx, My, size):
    """IceGrid.cc:443-484: processor grid Nx x Ny for `size` ranks."""
    if My <= 0:
        raise ValueError("'My' is invalid.")
    Nx = int(0.5 + math.sqrt(float(Mx) * float(size) / float(My)))
    Ny = 0
    if Nx == 0:
        Nx = 1
    while Nx > 0:
        Ny = size // Nx
        if Nx * Ny == size:
            break
        Nx -= 1
    if Mx > My and Nx < Ny:
        Nx, Ny = Ny, Nx
    if Mx // Nx < 2:
        raise ValueError("Can't split %d grid points into %d parts (X-direction)." % (Mx, Nx))
    if My // Ny < 2:
        raise ValueError("Can't split %d grid points into %d parts (Y-direction)." % (My, Ny))
    return Nx, Ny


def ownership_ranges(M, N):
    """IceGrid.cc:489-499."""
    return [M // N + (1 if (M % N) > i else 0) for i in range(N)]


@dataclass
class Patch:
    """One rank's owned rectangle (DMDA corners) and its neighbours in the periodic process grid."""
    rank: int
    px: int
    py: int
    Nx: int
    Ny: int
    xs: int
    xm: int
    ys: int
    ym: int

    def neighbor(self, dx, dy):
        """Rank of the periodic neighbour (DMDA is always periodic, IceGrid.cc:870-872)."""
        return ((self.px + dx) % self.Nx) + self.Nx * ((self.py + dy) % self.Ny)


def decompose(Mx, My, size, Nx=None, Ny=None, procs_x=None, procs_y=None):
    """All patches of a `size`-rank run, rank = px + Nx * py (DMDA ordering).

    Nx, Ny: PISM's -Nx / -Ny (default: compute_nprocs).  procs_x, procs_y: PISM's -procs_x / -procs_y, explicit
    ownership ranges (IceGrid.cc:519-586; they must have Nx / Ny entries and sum to Mx / My, :1327-1333); default:
    the uniform ranges of ownership_ranges()."""
    if Nx is None or Ny is None:
        Nx, Ny = compute_nprocs(Mx, My, size)
    if Nx * Ny != size:
        raise ValueError("Nx * Ny has to be equal to %d." % size)
    lx = list(procs_x) if procs_x is not None else ownership_ranges(Mx, Nx)
    ly = list(procs_y) if procs_y is not None else ownership_ranges(My, Ny)
    if len(lx) != Nx:
        raise ValueError("-Nx has to be equal to the -procs_x size.")
    if len(ly) != Ny:
        raise ValueError("-Ny has to be equal to the -procs_y size.")
    if sum(lx) != Mx:
        raise ValueError("procs_x don't sum up to Mx")
    if sum(ly) != My:
        raise ValueError("procs_y don't sum up to My")
    x0 = np.concatenate([[0], np.cumsum(lx)])
    y0 = np.concatenate([[0], np.cumsum(ly)])
    out = []
    for py in range(Ny):
        for px in range(Nx):
            out.append(Patch(px + Nx * py, px, py, Nx, Ny, int(x0[px]), int(lx[px]), int(y0[py]), int(ly[py])))
    return out


def _min_bottleneck_cuts(block_costs, parts, min_len=2):
    """Cut the rows of block_costs[n_rows, n_blocks] into `parts` consecutive groups so that the largest
    (group, block) sum is as small as possible: bisection on the bottleneck, greedy sweep for feasibility."""
    n = block_costs.shape[0]
    cum = np.vstack([np.zeros((1, block_costs.shape[1])), np.cumsum(block_costs, axis=0)])

    def sweep(limit):
        cuts, start = [], 0
        for g in range(parts):
            remaining = parts - g - 1
            lo, hi = start + min_len, n - remaining * min_len
            if g == parts - 1:
                end = n
            else:
                # the furthest end whose block sums stay below the limit (sums are monotone in end)
                seg = (cum[lo:hi + 1] - cum[start]).max(axis=1)
                ok = np.nonzero(seg <= limit)[0]
                end = lo + (int(ok[-1]) if len(ok) else 0)
            if (cum[end] - cum[start]).max() > limit:
                return None
            cuts.append(end - start)
            start = end
        return cuts

    lo, hi = float(block_costs.sum(axis=0).max()) / parts, float(block_costs.sum(axis=0).max())
    best = sweep(hi)
    for _ in range(50):
        mid = 0.5 * (lo + hi)
        c = sweep(mid)
        if c is None:
            lo = mid
        else:
            best, hi = c, mid
    return best


def balanced_ownership_ranges(cost, Nx, Ny, sweeps=3):
    """Ownership ranges (procs_x, procs_y) of an Nx x Ny tensor-product decomposition that balance a per-column cost
    map cost[My, Mx] (e.g. 1 for an ice-free column, ~2.5 for an icy one): what a PISM user passes as -procs_x /
    -procs_y.  Alternates 1D minimum-bottleneck cuts in y (given the x blocks) and in x (given the y blocks)."""
    My, Mx = cost.shape
    lx, ly = ownership_ranges(Mx, Nx), ownership_ranges(My, Ny)
    for _ in range(sweeps):
        xe = np.concatenate([[0], np.cumsum(lx)])
        rows_by_xblock = np.stack([cost[:, xe[b]:xe[b + 1]].sum(axis=1) for b in range(Nx)], axis=1)
        ly = _min_bottleneck_cuts(rows_by_xblock, Ny)
        ye = np.concatenate([[0], np.cumsum(ly)])
        cols_by_yblock = np.stack([cost[ye[b]:ye[b + 1], :].sum(axis=0) for b in range(Ny)], axis=1)
        lx = _min_bottleneck_cuts(cols_by_yblock, Nx)
    return [int(v) for v in lx], [int(v) for v in ly]


class Grid:
    """Computational box [-Lx, Lx] x [-Ly, Ly] x [0, Lz], cell-corner registration (IceGrid.cc:588-659)."""

    def __init__(self, Mx, My, Mz, Lx, Ly, Lz, spacing="equal", x0=0.0, y0=0.0, z=None):
        self.Mx, self.My, self.Mz = int(Mx), int(My), int(Mz)
        self.Lx, self.Ly, self.Lz = float(Lx), float(Ly), float(Lz)
        self.dx = 2.0 * self.Lx / (self.Mx - 1)
        self.dy = 2.0 * self.Ly / (self.My - 1)
        self.x = (x0 - self.Lx) + np.arange(self.Mx, dtype=np.float64) * self.dx
        self.x[-1] = x0 + self.Lx
        self.y = (y0 - self.Ly) + np.arange(self.My, dtype=np.float64) * self.dy
        self.y[-1] = y0 + self.Ly
        self.z = compute_vertical_levels(Lz, Mz, spacing) if z is None else np.asarray(z, dtype=np.float64)

    def whole(self):
        return Patch(0, 0, 0, 1, 1, 0, self.Mx, 0, self.My)

    def k_below_height(self, height):
        """IceGrid::kBelowHeight, IceGrid.cc:427-440 (GSL bsearch semantics: result in [0, Mz-2])."""
        if height < 0.0 - 1.0e-6:
            raise RuntimeError("height = %5.4f is below base of ice (height must be non-negative)" % height)
        if height > self.Lz + 1.0e-6:
            raise RuntimeError("height = %5.4f is above top of computational grid Lz = %5.4f" % (height, self.Lz))
        k = int(np.searchsorted(self.z, height, side="right")) - 1
        return min(max(k, 0), self.Mz - 2)


def local_shape(patch, w, dof=1):
    s = (patch.ym + 2 * w, patch.xm + 2 * w)
    return s if dof == 1 else s + (dof,)


def wrap_ghosts(a, w):
    """Periodic self-wrap of a whole-domain local array a[j, i, ...] with ghost width w, in place."""
    if w == 0:
        return a
    a[:, :w] = a[:, -2 * w:-w]
    a[:, -w:] = a[:, w:2 * w]
    a[:w, :] = a[-2 * w:-w, :]
    a[-w:, :] = a[w:2 * w, :]
    return a


def global_to_local(g, patch, w):
    """Cut a patch (+ periodic ghosts of width w) out of a global array g[j, i, ...]."""
    My, Mx = g.shape[0], g.shape[1]
    jj = np.arange(patch.ys - w, patch.ys + patch.ym + w) % My
    ii = np.arange(patch.xs - w, patch.xs + patch.xm + w) % Mx
    return np.ascontiguousarray(g[np.ix_(jj, ii)])
