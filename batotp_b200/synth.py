"""Seeded synthetic spline paths for the batch configurations.

Recipes follow the reference's own Octave generators, with their unseeded ``rand``
replaced by a counter-based splitmix64 stream so that every path is reproducible
from ``(config_id, path_index)`` alone:

* GEN7DOF  — input/GEN7DOF/generateGEN7DOFpath.m:7-10   20 knots U[0,5]^7,
             not-a-knot spline ('spline' in interp1), 400 points, tres 0.01
* CSPR3DOF — input/CSPR3DOF/generatePathPointsCSPR.m:5-23  20 knots x=3(u-.5),
             y=3(u-.35), z=3(u+.75), ss=0:0.005:19 -> 3801 points, Cartesian only.
             Reject-and-redraw: a candidate is kept only if it stays inside the STATIC
             workspace of the robot, i.e. the cable tensions that hold the platform at rest,
             A(x) tau = (0,0,g) (robot.cpp:534-558, 507-515), stay within [1.05, 11.5] N (the limits of
             input/CSPR3DOF/config.dat are [1, 12] N) at every 10th path point.  Inside the static
             workspace sdot -> 0, sddot = 0 is always feasible, so the bisection of ba.cpp:1248-1332 cannot
             fail and the reference never crawls to maxIntegTime; a rejected path index draws its next
             candidate from the stream (config_id + 256*attempt), so every path stays addressable by index
* KUKA     — SURVEY §8d C3: 20 knots, joint j ~ U[-0.8,0.8]*limit_j deg with limits
             (170,120,170,120,170,120,170), 400 points, tres 0.5 (as KUKApath.dat)

Everything is elementwise float64 numpy arithmetic (own tridiagonal solve, no LAPACK),
so the float32 payload bytes are identical on every machine.  The payload layout is
the BIN file's: [path][coordinate][point] float32 (ba.cpp:2283-2299).
"""
from __future__ import annotations

import numpy as np

CONFIG_IDS = {"GEN7DOF": 5, "CSPR3DOF": 4, "KUKA": 3}
_GAMMA = np.uint64(0x9E3779B97F4A7C15)


def splitmix_uniform(config_id, path_index: np.ndarray, n_draws: int) -> np.ndarray:
    """U[0,1) doubles, shape [len(path_index), n_draws]; seed = 0x5EED0000 + config_id*2^32 + index
    (config_id: int, or one per path)."""
    with np.errstate(over="ignore"):
        cid = np.broadcast_to(np.asarray(config_id, dtype=np.uint64), path_index.shape)
        seed = (np.uint64(0x5EED0000) + (cid << np.uint64(32)) + path_index.astype(np.uint64))[:, None]
        k = (np.arange(1, n_draws + 1, dtype=np.uint64))[None, :]
        z = seed + k * _GAMMA
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        z = z ^ (z >> np.uint64(31))
    return (z >> np.uint64(11)).astype(np.float64) * (1.0 / 9007199254740992.0)


def notaknot_eval(y: np.ndarray, s_out: np.ndarray) -> np.ndarray:
    """Not-a-knot cubic spline through y[..., n] at knots 0..n-1, evaluated at s_out -> [..., len(s_out)]."""
    n = y.shape[-1]
    rhs = 6.0 * (y[..., :-2] - 2.0 * y[..., 1:-1] + y[..., 2:])  # rows 1..n-2
    M = np.zeros(y.shape, dtype=np.float64)
    # not-a-knot: M0 = 2M1 - M2 and M_{n-1} = 2M_{n-2} - M_{n-3}  =>  6*M1 = rhs_1, 6*M_{n-2} = rhs_{n-2}
    M[..., 1] = rhs[..., 0] / 6.0
    M[..., n - 2] = rhs[..., n - 3] / 6.0
    # interior rows 2..n-3: M_{i-1} + 4 M_i + M_{i+1} = rhs_i with M_1, M_{n-2} known (Thomas)
    m = n - 4
    if m > 0:
        d = rhs[..., 1:n - 3].copy()
        d[..., 0] -= M[..., 1]
        d[..., m - 1] -= M[..., n - 2]
        cp = np.zeros(m)
        cp[0] = 1.0 / 4.0
        d[..., 0] = d[..., 0] / 4.0
        for i in range(1, m):
            den = 4.0 - cp[i - 1]
            cp[i] = 1.0 / den
            d[..., i] = (d[..., i] - d[..., i - 1]) / den
        for i in range(m - 2, -1, -1):
            d[..., i] = d[..., i] - cp[i] * d[..., i + 1]
        M[..., 2:n - 2] = d
    M[..., 0] = 2.0 * M[..., 1] - M[..., 2]
    M[..., n - 1] = 2.0 * M[..., n - 2] - M[..., n - 3]
    seg = np.minimum(np.floor(s_out).astype(np.int64), n - 2)
    t = s_out - seg
    y0 = y[..., seg]
    y1 = y[..., seg + 1]
    m0 = M[..., seg]
    m1 = M[..., seg + 1]
    c1 = (y1 - y0) - (m1 + 2.0 * m0) / 6.0
    c2 = m0 / 2.0
    c3 = (m1 - m0) / 6.0
    return np.ascontiguousarray(y0 + t * (c1 + t * (c2 + t * c3)))


def gen7dof_paths(first: int, count: int, n_knots: int = 20, n_pts: int = 400):
    """-> (tres, theta f32 [count, 7, n_pts])."""
    idx = np.arange(first, first + count)
    u = splitmix_uniform(CONFIG_IDS["GEN7DOF"], idx, n_knots * 7).reshape(count, n_knots, 7)
    knots = np.ascontiguousarray((5.0 * u).transpose(0, 2, 1))  # [count, 7, n_knots]
    s_out = np.linspace(0.0, n_knots - 1, n_pts)
    return float(np.float32(0.01)), np.ascontiguousarray(notaknot_eval(knots, s_out).astype(np.float32))


def kuka_paths(first: int, count: int, n_knots: int = 20, n_pts: int = 400):
    """-> (tres, theta f32 [count, 7, n_pts]) in degrees."""
    idx = np.arange(first, first + count)
    u = splitmix_uniform(CONFIG_IDS["KUKA"], idx, n_knots * 7).reshape(count, n_knots, 7)
    lim = np.array([170.0, 120.0, 170.0, 120.0, 170.0, 120.0, 170.0])
    knots = np.ascontiguousarray(((1.6 * u - 0.8) * lim).transpose(0, 2, 1))
    s_out = np.linspace(0.0, n_knots - 1, n_pts)
    return 0.5, np.ascontiguousarray(notaknot_eval(knots, s_out).astype(np.float32))


def cspr_pmat() -> np.ndarray:
    """Cable attachment points p[coordinate][cable] (robot.cpp:291-322)."""
    cible1 = np.array([1.0941, -4.9074, 2.5542])
    delta1 = np.array([-0.765, 0.112, 3.74])
    cible3 = np.array([0.2098, 5.3409, 2.6236])
    delta2 = np.array([0.43, 0.125, 3.615])
    p1, p2, p3 = cible1 + delta1, cible3 + delta2, np.array([-5.9751, 0.1399, 6.1543])
    pm = np.zeros((3, 3))
    for i, it in enumerate((1, 0, 2)):
        pm[i] = (-p1[it], -p2[it], -p3[it])
    return pm - (pm.sum(axis=1) / 3.0)[:, None]


def cspr_static_tensions(cart: np.ndarray) -> np.ndarray:
    """Cable tensions that hold the platform at rest at every point of cart[..., 3, n] -> [..., n, 3]:
    A tau = (0, 0, g) with A[r][c] = (x_r - p_rc) / rho_c (robot.cpp:534-558), by Cramer's rule."""
    x = np.asarray(cart, dtype=np.float64)
    d = x[..., :, None, :] - cspr_pmat()[:, :, None]          # [..., r, c, n]
    A = d / np.sqrt((d * d).sum(axis=-3, keepdims=True))        # columns are unit vectors along the cables
    a, b, c = A[..., 0, :, :], A[..., 1, :, :], A[..., 2, :, :]  # rows, each [..., cable, n]
    # b = (0,0,g): tau_k = g * cofactor(2,k) / det
    c0 = a[..., 1, :] * b[..., 2, :] - a[..., 2, :] * b[..., 1, :]
    c1 = a[..., 2, :] * b[..., 0, :] - a[..., 0, :] * b[..., 2, :]
    c2 = a[..., 0, :] * b[..., 1, :] - a[..., 1, :] * b[..., 0, :]
    det = c[..., 0, :] * c0 + c[..., 1, :] * c1 + c[..., 2, :] * c2
    with np.errstate(divide="ignore", invalid="ignore"):
        return np.stack([9.81 * c0 / det, 9.81 * c1 / det, 9.81 * c2 / det], axis=-1)


CSPR_STATIC_RANGE = (1.05, 11.5)  # N; config limits are [1, 12]


def cspr_accept(cart_f32: np.ndarray, stride: int = 10) -> np.ndarray:
    """-> bool[count]: the path stays inside the static workspace (every `stride`-th point and the last one)."""
    n = cart_f32.shape[-1]
    idx = np.unique(np.concatenate([np.arange(0, n, stride), [n - 1]]))
    tau = cspr_static_tensions(cart_f32[..., idx])
    lo, hi = CSPR_STATIC_RANGE
    with np.errstate(invalid="ignore"):
        return np.isfinite(tau).all(axis=(-1, -2)) & (tau.min(axis=(-1, -2)) >= lo) & (tau.max(axis=(-1, -2)) <= hi)


def cspr_paths(first: int, count: int, n_knots: int = 20, sres: float = 0.005, redraw: bool = True):
    """-> (tres, cart f32 [count, 3, n_pts]) ; n_pts = 3801 for the stock sres.  redraw=False returns the raw
    candidates of attempt 0 (some of which leave the static workspace; the reference's bisection fails there)."""
    n_pts = int(round((n_knots - 1) / sres)) + 1
    s_out = np.arange(n_pts, dtype=np.float64) * sres
    off = np.array([-0.5, -0.35, 0.75])
    out = np.empty((count, 3, n_pts), dtype=np.float32)
    pending = np.arange(count)
    for attempt in range(64):
        idx = first + pending
        u = splitmix_uniform(CONFIG_IDS["CSPR3DOF"] + 256 * attempt, idx, n_knots * 3).reshape(len(idx), n_knots, 3)
        knots = np.ascontiguousarray((3.0 * (u + off)).transpose(0, 2, 1))
        cand = notaknot_eval(knots, s_out).astype(np.float32)
        ok = cspr_accept(cand) if redraw else np.ones(len(idx), dtype=bool)
        out[pending[ok]] = cand[ok]
        pending = pending[~ok]
        if len(pending) == 0:
            break
    else:
        raise RuntimeError("cspr_paths: no candidate inside the static workspace after 64 draws")
    return float(np.float32(sres)), out
