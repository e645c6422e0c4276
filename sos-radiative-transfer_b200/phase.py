"""Azimuth-averaged phase functions P0(mu, mu0) and P(mu, mu') for the analytic / tabulated families.

Input producer of the hot path.  Same quadrature as the
reference builders -- 25 azimuth nodes on [0, pi], both half-rings, composite trapezoid, raw
matrix symmetric, then every *column* normalised so that trapz(P[:, n], mu) = 4
(SOS_Aer_phase_func.py:68-292) -- but evaluated as array expressions instead of the reference's
triple Python loop (89-112 s per matrix at N = 1002 there, well under a second here).

Log-normal Mie aerosols (SOS_Aer_phase_func.py:398-753, the reference's 'eva' / 'wildfire' families) need
`miepython`, which the reference does not pin and which is not installed here: family 'mie_lognormal' takes
its tabulated mixture phase function from the host Lorenz-Mie stand-in of mie.py (g = (wavelength, n_re, n_im,
r_m, sigma), e.g. mie.EVA_AEROSOL) and then goes through the same tabulated builder as the FWC cloud.  Callers
with their own Mie code still pass (P0, P) arrays.
"""
from __future__ import annotations

import os

import numpy as np

NB_PHI = 25  # SOS_Aer_phase_func.py:81

_FWC = None


def _fwc_table():
    """The tabulated FWC cloud phase function (data of SOS_Aer_fwc_data.py:3,173)."""
    global _FWC
    if _FWC is None:
        path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "data", "fwc_table.npz")
        d = np.load(path)
        _FWC = (d["mu_fwc"], d["phase_func_FWC"])
    return _FWC


def phase_table(name, g=None):
    """(cos Theta ascending, values) of a tabulated family: 'fwc', or 'mie_lognormal' with g = (wl, n_re, n_im, r_m, sigma)."""
    name, g = resolve(name, g)
    if name == "fwc":
        return _fwc_table()
    if name == "mie_lognormal":
        from . import mie
        return mie.lognormal_table(*[float(v) for v in g])
    raise ValueError(f"{name!r} is not a tabulated phase function")


TABULATED = ("fwc", "mie_lognormal")


def resolve(name, g=None):
    """The reference's dispatcher names for its two Mie aerosols (phase_func(), SOS_Aer_phase_func.py:12-63: 'eva',
    'wildfire') -> ('mie_lognormal', parameters); every other family is returned unchanged."""
    if name in ("eva", "wildfire"):
        from . import mie
        return "mie_lognormal", (mie.EVA_AEROSOL if name == "eva" else mie.WILDFIRE_AEROSOL)
    return name, g


def _trapz(y, x, axis=-1):
    y = np.asarray(y)
    d = np.diff(x)
    y = np.moveaxis(y, axis, -1)
    return (d * (y[..., 1:] + y[..., :-1]) / 2.0).sum(-1)


def _kernel(name, g):
    if name == "rayleigh":
        return lambda c: 0.75 * (1 + c * c)  # SOS_Aer_phase_func.py:97
    if name == "hg":
        return lambda c: (1 - g * g) / ((1 + g * g - 2 * g * c) ** 1.5)  # :158
    if name in TABULATED:
        xs, ys = phase_table(name, g)

        def interp(c):  # interpolate_fwc_phase, :202-236
            c = np.clip(c, -1, 1)
            i = np.searchsorted(xs, c)
            lo = np.clip(i - 1, 0, len(xs) - 1)
            hi = np.clip(i, 0, len(xs) - 1)
            with np.errstate(all="ignore"):
                w = (c - xs[lo]) / (xs[hi] - xs[lo])
                v = ys[lo] + w * (ys[hi] - ys[lo])
            v = np.where(i == 0, ys[0], v)
            v = np.where(i >= len(xs), ys[-1], v)
            return v
        return interp
    raise ValueError(f"unknown analytic phase function {name!r}")


def phase_P0(name: str, nb_angles: int, mu: np.ndarray, mu0: float, g: float = 0.5) -> np.ndarray:
    """First-order phase function P0(mu, mu0), normalised to trapz(P0, mu) = 2 (:92-105)."""
    name, g = resolve(name, g)
    N = 2 * nb_angles
    mu = np.asarray(mu, dtype=np.float64)
    if name == "iso":
        return np.ones(N)  # :71
    f = _kernel(name, g)
    phi = np.linspace(0, np.pi, NB_PHI)
    cphi = np.cos(0 - phi)
    s = np.sqrt(1 - mu * mu)
    cc = (mu * mu0)[:, None]
    ss = (np.sqrt(1 - mu0 * mu0) * s)[:, None] * cphi[None, :]
    P0 = _trapz(f(-(cc + ss)) + f(-(cc - ss)), phi) / (4 * np.pi)
    return P0 / _trapz(P0, mu) * 2


def phase_P(name: str, nb_angles: int, mu: np.ndarray, g: float = 0.5, block: int = 128) -> np.ndarray:
    """P(mu, mu'): raw matrix symmetric, then every column normalised to trapz = 4 (:112-131)."""
    name, g = resolve(name, g)
    N = 2 * nb_angles
    mu = np.asarray(mu, dtype=np.float64)
    if name == "iso":
        return 2 * np.ones((N, N))  # :74
    f = _kernel(name, g)
    phi = np.linspace(0, np.pi, NB_PHI)
    cphi = np.cos(0 - phi)
    s = np.sqrt(1 - mu * mu)
    P = np.empty((N, N))
    for a in range(0, N, block):
        b = min(N, a + block)
        cmn = mu[:, None] * mu[None, a:b]                       # mu[m]*mu[n]
        smn = (s[None, a:b] * s[:, None])                       # sqrt(1-mu[n]^2)*sqrt(1-mu[m]^2)
        x = smn[:, :, None] * cphi[None, None, :]
        P[:, a:b] = _trapz(f(-(cmn[:, :, None] + x)) + f(-(cmn[:, :, None] - x)), phi) / (2 * np.pi)
    return 4 * P / _trapz(P, mu, axis=0)[None, :]


def phase_matrices(name: str, nb_angles: int, mu: np.ndarray, mu0: float, g: float = 0.5):
    """Return (P0 (N,), P (N, N)) for name in {'iso', 'rayleigh', 'hg', 'fwc', 'mie_lognormal'}."""
    return phase_P0(name, nb_angles, mu, mu0, g), phase_P(name, nb_angles, mu, g)
