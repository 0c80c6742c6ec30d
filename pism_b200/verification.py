"""Exact solutions used by pismv tests B, C, L and F/G, restated in numpy.

Reference sources (compiled unmodified into oracle/_ref/libpism_exact.so where the reference
tree is mounted; tests/test_exact_solutions.py checks these restatements against that build):
  * exactC:  src/verification/tests/exactTestsABCD.c:94-121
  * exactFG: src/verification/tests/exactTestsFG.cc:41-205
plus the set-ups that turn them into SIAFD inputs:
  * Test C:  src/verification/iceCompModel.cc:301-356, src/pismv.cc:96-102
  * Test F/G: src/verification/iCMthermo.cc:101-136, src/stressbalance/sia/siafd_test.cc:154-224
"""
import numpy as np

SperA = 31556926.0  # exactTestsABCD.c:26, exactTestsFG.cc:43


def exactC(t, r):
    """H(t, r), M(t, r) of Test C (exactTestsABCD.c:94-121). t in seconds, r array in metres."""
    r = np.asarray(r, dtype=np.float64)
    n, H0, R0 = 3.0, 3600.0, 750000.0
    lam, alpha, beta = 5.0, -1.0, 2.0
    t0 = 15208.0 * SperA
    t = np.float64(t)  # t = 0: pow(0, -beta) = inf as in C, and r < Rmargin = 0 is false everywhere
    Rmargin = R0 * (t / t0) ** beta
    with np.errstate(all="ignore"):
        inner = 1.0 - ((t / t0) ** (-beta) * (r / R0)) ** ((n + 1) / n)
        H = np.where(r < Rmargin, H0 * (t / t0) ** (-alpha) * np.where(inner > 0, inner, 0.0) ** (n / (2 * n + 1)), 0.0)
    if t > 0.1 * SperA:
        M = (lam / t) * H
    else:
        Rm = R0 * (0.1 * SperA / t0) ** beta
        M = np.where(r < Rm, 5 * H0 / t0, 0.0)
    return H, M


def exactB(t, r):
    """H(t, r), M = 0 of Test B, the Halfar solution (exactTestsABCD.c:63-85). t in seconds."""
    r = np.asarray(r, dtype=np.float64)
    n, H0, R0 = 3.0, 3600.0, 750000.0
    alpha, beta = 1.0 / 9.0, 1.0 / 18.0
    t0 = 422.45 * SperA
    t = np.float64(t)
    Rmargin = R0 * (t / t0) ** beta
    with np.errstate(all="ignore"):
        inner = 1.0 - ((t / t0) ** (-beta) * (r / R0)) ** ((n + 1) / n)
        H = np.where(r < Rmargin, H0 * (t / t0) ** (-alpha) * np.where(inner > 0, inner, 0.0) ** (n / (2 * n + 1)), 0.0)
    return H, np.zeros_like(H)


def _p3(x):
    # exactTestsFG.cc:30-33
    return -6.0 + x * (6.0 + x * (-3.0 + x))


def exactFG(t, r, z, Cp):
    """Test F (Cp = 0) / G exact solution at one radius r (0 < r < L) on levels z.

    Returns dict(H, T[Mz], U[Mz]) -- the quantities SIAFD inputs and checks need
    (exactTestsFG.cc:41-130); w, Sig, Sigc are not restated (out of this path's scope).
    """
    z = np.asarray(z, dtype=np.float64)
    H0, L = 3000.0, 750000.0
    Tp = 2000.0 * SperA
    g, Rgas = 9.81, 8.314
    rho, k, n = 910.0, 2.1, 3.0
    A, Q = 3.615e-13, 6.0e4
    Ggeo, ST, Tmin = 0.042, 1.67e-5, 223.15
    if r <= 0 or r >= L:
        raise ValueError("exactFG(): code and derivation assume 0 < r < L  !")
    power = n / (2 * n + 2)
    Hconst = H0 / (1 - 1 / n) ** power
    s = r / L
    lamhat = (1 + 1 / n) * s - (1 / n) + (1 - s) ** (1 + 1 / n) - s ** (1 + 1 / n)
    if 0.3 * L < r < 0.9 * L:
        f = np.cos(np.pi * (r - 0.6 * L) / (0.6 * L)) ** 2.0
    else:
        f = 0.0
    goft = Cp * np.sin(2.0 * np.pi * t / Tp)
    H = Hconst * lamhat ** power + goft * f
    Ts = Tmin + ST * r
    nusqrt = np.sqrt(1 + (4.0 * H * Ggeo) / (k * Ts))
    nu = (k * Ts / (2.0 * Ggeo)) * (1 + nusqrt)
    T = np.where(z < H, Ts * (nu + H) / (nu + z), Ts)
    lamhatr = ((1 + 1 / n) / L) * (1 - (1 - s) ** (1 / n) - s ** (1 / n))
    if 0.3 * L < r < 0.9 * L:
        fr = -(np.pi / (0.6 * L)) * np.sin(2.0 * np.pi * (r - 0.6 * L) / (0.6 * L))
    else:
        fr = 0.0
    Hr = Hconst * power * lamhat ** (power - 1) * lamhatr + goft * fr
    if Hr > 0:
        raise ValueError("exactFG(): assumes H_r negative for all 0 < r < L !")
    mu = Q / (Rgas * Ts * (nu + H))
    surfArr = np.exp(-Q / (Rgas * Ts))
    Uconst = 2.0 * (rho * g) ** n * A
    omega = Uconst * (-Hr) ** n * surfArr * mu ** (-n - 1)
    I3 = np.where(z < H, _p3(mu * H) * np.exp(mu * H) - _p3(mu * (H - z)) * np.exp(mu * (H - z)),
                  _p3(mu * H) * np.exp(mu * H) - _p3(0.0))
    U = omega * I3
    return {"H": float(H), "T": T, "U": U}


def exactL(r):
    """H(r), b(r), a(r) of Test L, the steady isothermal sheet on a non-flat bed
    (src/verification/tests/exactTestL.cc:35-178): u = H^(8/3) solves
        du/dr = -(8/3) b'(r) u^(5/8) - (a0 r (L^2 - r^2) / (2 L^2 Gamma~))^(1/3),   u(L) = 0,
    integrated inward from the margin.  The reference integrates with GSL's rk8pd at EPS_ABS = 1e-12 (GSL is not in
    this image: `gsl_odeiv2` cannot be built here); this restatement uses scipy's DOP853 at tolerances tight enough
    that the two agree far below the six decimals of the golden rows of test/regression/test_16.sh, which is what
    pins it (tests/test_oracle_pismv.py).  r: array of any shape, metres."""
    from scipy.integrate import solve_ivp
    r = np.asarray(r, dtype=np.float64)
    L, b0, z0, g, rho, n = 750.0e3, 500.0, 1.2, 9.81, 910.0, 3.0
    Lsqr = L * L
    a0 = 0.3 / SperA
    A = 1.0e-16 / SperA
    Gamma = 2 * (rho * g) ** n * A / (n + 2)
    tilGamma = Gamma * n ** n / (2.0 * n + 2.0) ** n
    Cc = a0 / (2.0 * Lsqr * tilGamma)
    freq = z0 * np.pi / L

    def funcL(rr, u):
        if 0.0 <= rr <= L:
            bprime = b0 * freq * np.sin(freq * rr)
            return [-(8.0 / 3.0) * bprime * max(u[0], 0.0) ** (5.0 / 8.0) - (Cc * rr * (Lsqr - rr * rr)) ** (1.0 / 3.0)]
        return [0.0]

    flat = r.ravel()
    inside = np.unique(flat[flat < L])[::-1]  # decreasing radii, like the sorted list of iceCompModel.cc:372-401
    u_of = {}
    if inside.size:
        sol = solve_ivp(funcL, (L, float(inside[-1])), [0.0], method="DOP853", t_eval=inside, rtol=1e-13, atol=1e-10,
                        first_step=1.0)
        assert sol.success, sol.message
        u_of = dict(zip(inside.tolist(), sol.y[0].tolist()))
    u = np.array([u_of.get(x, 0.0) for x in flat.tolist()]).reshape(r.shape)
    H = np.maximum(u, 0.0) ** (3.0 / 8.0)
    b = -b0 * np.cos(z0 * np.pi * r / L)
    a = a0 * (1.0 - (2.0 * r * r / Lsqr))
    return H, b, a


def radius(grid):
    """radius(grid, i, j) = sqrt(x^2 + y^2) (src/util/IceGrid.cc `radius`), as a [My, Mx] array."""
    X, Y = np.meshgrid(grid.x, grid.y)
    return np.sqrt(X * X + Y * Y)
