"""Host-side grid helpers: mu grid, tau profile, extrapolation tables.

These are the cheap, once-per-plan pieces of the reference that stay on the host
(SURVEY.md 2, rows #3 and #7); every heavy loop runs in libsos_b200.
"""
from __future__ import annotations

import functools

import numpy as np

# SOS_Aer_global_va.py:5-7
MU_THRESHOLD = 0.01
MU_EXTREME_THRESHOLD = 1e-8
MU_VERY_SMALL_THRESHOLD = 0.001


def mu_grid(nb_angles: int) -> np.ndarray:
    """mu = [linspace(-1,0,M), linspace(0,1,M)] (SOS_Aer_main_specular.py:59-61)."""
    return np.concatenate((np.linspace(-1, 0, nb_angles), np.linspace(0, 1, nb_angles)))


@functools.lru_cache(maxsize=256)
def _aerosol_rows_cached(z0, z_up, z_down, nb_layers):
    if z_down > z_up:
        z_down, z_up = z_up, z_down
    z = np.linspace(z0, 0, nb_layers)
    z.setflags(write=False)  # shared between callers
    return z, int(np.argmin(np.abs(z - z_up))), int(np.argmin(np.abs(z - z_down)))


def aerosol_rows(z0, z_up, z_down, nb_layers):
    """Layer indices bounding the aerosol layer (SOS_Aer_main_specular.py:30,39-40).  Returns
    (z_profile (read-only), idx_up, idx_down); cached, because a sweep asks thousands of times."""
    return _aerosol_rows_cached(float(z0), float(z_up), float(z_down), int(nb_layers))


def tau_profile(tauStar_atm, tauStar_aer, z0, z_up, z_down, nb_layers):
    """Cumulative optical depth: molecular ramp + aerosol ramp (SOS_Aer_tau_profile.py:5-53).

    Same values as the reference; the plotting side effect (Q20) is dropped.
    """
    _, idx_up, idx_down = aerosol_rows(z0, z_up, z_down, nb_layers)
    rows = np.arange(nb_layers)
    tau = rows * tauStar_atm / (nb_layers - 1)
    step = tauStar_aer / (idx_down + 1 - idx_up)
    inside = (rows >= idx_up) & (rows <= idx_down)
    tau[inside] += (rows[inside] + 1 - idx_up) * step
    tau[rows > idx_down] += tauStar_aer
    return tau


def mu_approx_In(mu, nb_angles):
    """Indices of the first mu >= 0.009 and >= 0.020 on the upward half (SOS_Aer_I1_In.py:274-282)."""
    up = np.asarray(mu)[nb_angles:]
    i1 = int(np.argmax(~(up < 0.009)))
    rest = ~(up[i1:] < 0.020)
    i2 = i1 + int(np.argmax(rest))
    if not (~(up < 0.009)).any() or not rest.any():
        raise IndexError("mu_approx_In: threshold not reached on the mu grid")
    return nb_angles + i1, nb_angles + i2


_WIDTH_FACTORS = (0.005, 0.02, 0.04, 0.06)


def extrapolation_width(tau_ref: float, nb_angles: int) -> int:
    """How many columns next to mu=0- are extrapolated (SOS_Aer_main_specular.py:342-345)."""
    if tau_ref <= 0.0625:
        f = _WIDTH_FACTORS[0]
    elif tau_ref <= 1:
        f = _WIDTH_FACTORS[1]
    elif tau_ref < 4:
        f = _WIDTH_FACTORS[2]
    else:
        f = _WIDTH_FACTORS[3]
    return int(f * nb_angles)


def extrapolation_widths(tau_ref, nb_angles: int) -> np.ndarray:
    """extrapolation_width for an array of reference optical depths (same thresholds, same truncation)."""
    t = np.asarray(tau_ref, dtype=np.float64)
    f = np.select([t <= 0.0625, t <= 1, t < 4], _WIDTH_FACTORS[:3], default=_WIDTH_FACTORS[3])
    return (f * nb_angles).astype(np.int32)   # (positive products: the same truncation as int(f * nb_angles))


def _extrapolation_weights(mu_down: np.ndarray, width: int) -> np.ndarray:
    """Weights W[i, j] with target column M-1-i = sum_j W[i,j] * source column j.

    The reference extrapolates towards mu=0- with a least-squares parabola through the
    min(5, width) columns just outside the extrapolated zone (np.polyfit degree 2), a straight
    line when only two are available, and a two-point line through columns M-3, M-2 when
    width == 1 (SOS_Aer_In_limit.py:113-141).  All three are linear in the source values, so they
    collapse into one small matrix per width.
    """
    M = len(mu_down)
    if width <= 0:
        return np.zeros((0, 0))
    targets = mu_down[M - 1 - np.arange(width)]
    if width == 1:
        x = mu_down[M - 3: M - 1]
        # slope through (x0, y0), (x1, y1), anchored at x1
        t = (targets[:, None] - x[1]) / (x[0] - x[1])
        return np.concatenate((t, 1 - t), axis=1)
    ns = min(5, width)
    x = mu_down[M - width - ns: M - width]
    if ns == 2:
        t = (targets[:, None] - x[0]) / (x[1] - x[0])
        return np.concatenate((1 - t, t), axis=1)
    W = np.empty((width, ns))
    for j in range(ns):
        e = np.zeros(ns)
        e[j] = 1.0
        W[:, j] = np.polyval(np.polyfit(x, e, 2), targets)
    return W


_TABLE_CACHE = {}


def extrapolation_tables(mu: np.ndarray, nb_angles: int):
    """The four W matrices (one per width class) flattened in sos_extrap_layout() order (cached per grid)."""
    mu_down = np.ascontiguousarray(mu[:nb_angles], dtype=np.float64)
    key = (int(nb_angles), mu_down.tobytes())
    hit = _TABLE_CACHE.get(key)
    if hit is None:
        parts = []
        for f in _WIDTH_FACTORS:
            w = int(f * nb_angles)
            parts.append(_extrapolation_weights(mu_down, w).ravel())
        hit = np.concatenate(parts) if parts else np.zeros(0)
        hit.setflags(write=False)
        if len(_TABLE_CACHE) > 32:
            _TABLE_CACHE.clear()
        _TABLE_CACHE[key] = hit
    return hit
