"""Host Lorenz-Mie stand-in for the reference's `miepython` calls.

The reference builds its EVA / wildfire aerosol phase functions with `miepython` (unpinned, not installed
here): `miepython.efficiencies` for the scattering efficiency of every radius of a log-normal size
distribution (SOS_Aer_phase_func.py:419) and `miepython.i_unpolarized` for the unpolarised scattered
intensity (SOS_Aer_phase_func.py:334-335,362,693).  This module restates the published Lorenz-Mie series
(Bohren & Huffman 1983, ch. 4 and appendix A: logarithmic derivative D_n(mx) by downward recurrence,
Riccati-Bessel functions by upward recurrence, angular functions pi_n / tau_n by recurrence) with the
series cut at x + 4 x^(1/3) + 2 terms (Wiscombe 1980), vectorised over the scattering angles, and the
size-distribution mixing of `log_normal_mie` (SOS_Aer_phase_func.py:398-491,684-753) AS CODED, including its
two quirks (`as_coded=True`):
  * the weight of a radius is n(r) * Qsca without the geometric cross-section r^2 (:421), and
  * `efficiencies(idx, x_list, wl)` receives size parameters in the diameter slot (:419), i.e.
    Qsca is evaluated at x_eff = pi * x / wl.
`as_coded=False` gives the physical mixture (weights n(r) r^2 Qsca(x)).

Parity status: "Mie stand-in" -- the series is pinned on published values (tests/test_host_logic.py: Bohren &
Huffman's test sphere, Wiscombe's m = 1.5, x = 10 case, the Rayleigh limit, the optical theorem), not on
miepython itself.  The hot path does not depend on it: both implementations consume the same P0 / P arrays.

The result is a TABLE p(cos Theta) on the reference's 6001-point grid (compute_P, :684-694); the azimuth average
into P0(mu, mu0) / P(mu, mu') is the tabulated-family builder shared with the FWC cloud (host: phase.py, device:
sos_build_phase family 2), because interpolate_phase (:696-711) is the same linear interpolation as
interpolate_fwc_phase (:202-236) and mixing commutes with it.
"""
from __future__ import annotations

import functools
import hashlib
import os

import numpy as np

CACHE_VERSION = 1                     # bump when the table's definition changes
N_ANGLES_TABLE = 6001                 # compute_P, SOS_Aer_phase_func.py:685
RADII = (0.01, 10.0, 100)             # list_radius = linspace(0.01, 10, 100) micrometres, :404-405

# the two aerosol scenarios of the reference README (README.md:90-111): (wavelength um, n_re, n_im, r_m um, sigma)
EVA_AEROSOL = (0.55, 1.44, 0.0, 0.506, 1.2)
WILDFIRE_AEROSOL = (0.55, 1.7, 0.03, 0.065, 1.5)


def _coefficients(m: complex, x: float):
    """Mie coefficients a_n, b_n (n = 1..nmax) of a sphere of relative index m (Im m >= 0 absorbs) and size parameter x."""
    nmax = int(x + 4.05 * x ** (1.0 / 3.0) + 2.0)
    mx = m * x
    nmx = int(max(nmax, abs(mx)) + 16)
    D = np.zeros(nmx + 1, dtype=np.complex128)          # D_n(mx), downward recurrence
    for n in range(nmx, 0, -1):
        D[n - 1] = n / mx - 1.0 / (D[n] + n / mx)
    a = np.empty(nmax, dtype=np.complex128)
    b = np.empty(nmax, dtype=np.complex128)
    psi0, psi1 = np.cos(x), np.sin(x)                   # psi_{-1}, psi_0
    chi0, chi1 = -np.sin(x), np.cos(x)                  # chi_{-1}, chi_0
    for n in range(1, nmax + 1):
        psi = (2 * n - 1) / x * psi1 - psi0
        chi = (2 * n - 1) / x * chi1 - chi0
        xi, xi1 = psi - 1j * chi, psi1 - 1j * chi1
        da = D[n] / m + n / x
        db = m * D[n] + n / x
        a[n - 1] = (da * psi - psi1) / (da * xi - xi1)
        b[n - 1] = (db * psi - psi1) / (db * xi - xi1)
        psi0, psi1, chi0, chi1 = psi1, psi, chi1, chi
    return a, b


def efficiencies(m: complex, x: float):
    """(Qext, Qsca, Qback, g) of one sphere."""
    m = complex(m.real, abs(m.imag))
    a, b = _coefficients(m, x)
    n = np.arange(1, len(a) + 1)
    qext = 2.0 / x ** 2 * np.sum((2 * n + 1) * (a + b).real)
    qsca = 2.0 / x ** 2 * np.sum((2 * n + 1) * (np.abs(a) ** 2 + np.abs(b) ** 2))
    qback = np.abs(np.sum((2 * n + 1) * (-1.0) ** n * (a - b))) ** 2 / x ** 2
    gsum = np.sum(n[:-1] * (n[:-1] + 2) / (n[:-1] + 1) * (a[:-1] * np.conj(a[1:]) + b[:-1] * np.conj(b[1:])).real)
    gsum += np.sum((2 * n + 1) / (n * (n + 1)) * (a * np.conj(b)).real)
    return float(qext), float(qsca), float(qback), float(4.0 / x ** 2 * gsum / qsca)


def i_unpolarized(m: complex, x: float, mu) -> np.ndarray:
    """Unpolarised scattered intensity (|S1|^2 + |S2|^2) / 2 at mu = cos(scattering angle), normalised so that its
    integral over 4 pi steradians is the single-scattering albedo Qsca / Qext (miepython's default 'albedo' norm)."""
    m = complex(m.real, abs(m.imag))
    mu = np.atleast_1d(np.asarray(mu, dtype=np.float64))
    a, b = _coefficients(m, x)
    n = np.arange(1, len(a) + 1)
    qext = 2.0 / x ** 2 * np.sum((2 * n + 1) * (a + b).real)
    S1 = np.zeros(mu.shape, dtype=np.complex128)
    S2 = np.zeros(mu.shape, dtype=np.complex128)
    pi0 = np.zeros_like(mu)   # pi_{n-1}
    pi1 = np.ones_like(mu)    # pi_n, starting at n = 1
    for k in range(1, len(a) + 1):
        tau = k * mu * pi1 - (k + 1) * pi0
        f = (2 * k + 1) / (k * (k + 1))
        S1 += f * (a[k - 1] * pi1 + b[k - 1] * tau)
        S2 += f * (a[k - 1] * tau + b[k - 1] * pi1)
        pi0, pi1 = pi1, ((2 * k + 1) * mu * pi1 - (k + 1) * pi0) / k
    return (np.abs(S1) ** 2 + np.abs(S2) ** 2) / 2.0 / (np.pi * x ** 2 * qext)


def cache_dir() -> str:
    """Directory of the parameter-keyed table cache: $SOS_B200_CACHE, else ~/.cache/sos_b200."""
    return os.environ.get("SOS_B200_CACHE") or os.path.join(os.path.expanduser("~"), ".cache", "sos_b200")


def _table_path(wl, n_re, n_im, r_m, sig, as_coded) -> str:
    # Every parameter the table depends on is in the file name (the reference keys its .npy cache in the working
    # directory on nb_angles and a few scenario numbers only, so stale tables survive a change of refractive index:
    # SOS_Aer_global_va.py:17-83, SOS_Aer_phase_func.py:24-33 -- Q16), plus a format version.
    key = repr((CACHE_VERSION, float(wl), float(n_re), float(abs(n_im)), float(r_m), float(sig), bool(as_coded), RADII, N_ANGLES_TABLE))
    return os.path.join(cache_dir(), "mie_lognormal_" + hashlib.sha1(key.encode()).hexdigest()[:20] + ".npy")


@functools.lru_cache(maxsize=8)
def lognormal_table(wl: float, n_re: float, n_im: float, r_m: float, sig: float, as_coded: bool = True):
    """(cos Theta grid (6001,), mixture phase function on it) of a log-normal population of spheres.

    SOS_Aer_phase_func.py:404-421 (radii, n(r), Qsca weights), :684-694 (per-radius phase functions), :713-753
    (mixture = trapz over the radius of n(r) Qsca(r) p_r).  The overall scale is irrelevant: P0 and every column of P
    are normalised afterwards (:131).

    The table is cached in memory (per process) and on disk as a .npy file keyed on every parameter (cache_dir();
    SOS_B200_CACHE=off disables the disk cache): the Lorenz-Mie series over 200 radii x 6001 angles takes seconds."""
    mu_s = np.linspace(-1.0, 1.0, N_ANGLES_TABLE)
    mu_s.setflags(write=False)
    path = None if os.environ.get("SOS_B200_CACHE", "") == "off" else _table_path(wl, n_re, n_im, r_m, sig, as_coded)
    if path and os.path.exists(path):
        try:
            mix = np.load(path)
            if mix.shape == (N_ANGLES_TABLE,) and mix.dtype == np.float64 and np.all(np.isfinite(mix)):
                mix.setflags(write=False)
                return mu_s, mix
        except (OSError, ValueError):
            pass  # unreadable file: rebuild and overwrite
    mix = _lognormal_mixture(wl, n_re, n_im, r_m, sig, as_coded, mu_s)
    mix.setflags(write=False)
    if path:
        try:
            os.makedirs(os.path.dirname(path), exist_ok=True)
            tmp = path + ".%d.tmp" % os.getpid()
            with open(tmp, "wb") as f:
                np.save(f, mix)
            os.replace(tmp, path)   # atomic: concurrent ranks may build the same table
        except OSError:
            pass  # read-only location: the in-memory cache still works
    return mu_s, mix


def _lognormal_mixture(wl, n_re, n_im, r_m, sig, as_coded, mu_s):
    m = complex(n_re, abs(n_im))
    radii = np.linspace(*RADII)
    n_r = (1.0 / radii) * np.exp(-((np.log(radii) - np.log(r_m)) ** 2) / (2.0 * np.log(sig) ** 2))
    x_list = 2.0 * np.pi * radii / wl
    table = np.zeros((len(radii), N_ANGLES_TABLE))
    weight = np.empty(len(radii))
    for i, x in enumerate(x_list):
        if as_coded:
            weight[i] = n_r[i] * efficiencies(m, np.pi * x / wl)[1]      # Q23: x lands in the diameter slot
        else:
            weight[i] = n_r[i] * radii[i] ** 2 * efficiencies(m, x)[1]
        table[i] = i_unpolarized(m, x, mu_s)
    return _trapz_rows(weight[:, None] * table, radii)


def _trapz_rows(y, x):
    d = np.diff(x)
    return (d[:, None] * (y[1:] + y[:-1]) / 2.0).sum(0)
