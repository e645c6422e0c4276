"""CPU oracle for the SOS_AER hot path -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

A fresh NumPy restatement of the reference algorithm (Guillaume-SOULIER/
SOS-Radiative-Transfer, pure Python/NumPy).  Only tests/, __graft_entry__.smoke()
and bench.py's cpu_baseline / --impl reference leg may import this module; the
product package (sos-radiative-transfer_b200/) never does and fails loudly when
its CUDA library is missing.

Pinned: yes.  The reference ships no golden vectors or tests, so
the pin is the reference itself, executed unmodified in the build container
through oracle/ref_harness.py; tests/golden/*.npz hold its outputs and
tests/golden/make_golden.py is the script that produced them.
tests/test_oracle_vs_golden.py checks every function below against them.

Two evaluation methods are offered for the layer integration:
  method="slices"      the reference's own O(L^2 N) scheme: every (layer, mu)
                       value is a trapezoid over the whole slice above/below it
                       (SOS_Aer_I1_In.py:88-122, SOS_Aer_main_specular.py:330-449),
                       vectorised over mu instead of the reference's Python loop
                       over m.  This is the "port" timed as the CPU baseline.
  method="recurrence"  the algebraically identical O(L N) exp(-dtau/mu) linear
                       recurrence (rounding-level different, <1e-13), used to
                       check the GPU at sizes where "slices" is too slow.

All citations are file:line under /root/reference.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import List, Optional, Sequence, Tuple

import numpy as np

# SOS_Aer_global_va.py:5-7
MU_THRESHOLD = 0.01
MU_EXTREME_THRESHOLD = 1e-8
MU_VERY_SMALL_THRESHOLD = 0.001
BLEND_THRESHOLD = 0.0001  # SOS_Aer_I1_In.py:103
MU0_TOLERANCE = 0.0001  # SOS_Aer_I1_In.py:41, SOS_Aer_main_specular.py:111,204
CONVERGENCE = 0.0001  # SOS_Aer_main_specular.py:309


def trapz(y, x, axis=-1):
    """Composite trapezoid, same formula as np.trapz: sum(dx*(y[1:]+y[:-1])/2)."""
    y = np.asarray(y, dtype=np.float64)
    x = np.asarray(x, dtype=np.float64)
    d = np.diff(x)
    if y.ndim > 1:
        shape = [1] * y.ndim
        shape[axis] = d.shape[0]
        d = d.reshape(shape)
    sl1 = [slice(None)] * y.ndim
    sl0 = [slice(None)] * y.ndim
    sl1[axis] = slice(1, None)
    sl0[axis] = slice(None, -1)
    return (d * (y[tuple(sl1)] + y[tuple(sl0)]) / 2.0).sum(axis)


# --------------------------------------------------------------------------
# grid helpers
# --------------------------------------------------------------------------
def mu_grid(nb_angles: int) -> np.ndarray:
    """SOS_Aer_main_specular.py:59-61 (mu=0 appears twice)."""
    return np.concatenate((np.linspace(-1, 0, nb_angles), np.linspace(0, 1, nb_angles)))


def tau_profile(tauStar_atm, tauStar_aer, z0, z_up, z_down, nb_layers):
    """SOS_Aer_tau_profile.py:15-27 without the plotting side effect."""
    z = np.linspace(z0, 0, nb_layers)
    idx_up = int(np.argmin(np.abs(z - z_up)))
    idx_down = int(np.argmin(np.abs(z - z_down)))
    tau = np.arange(0, nb_layers) * tauStar_atm / (nb_layers - 1)
    step = tauStar_aer / (idx_down + 1 - idx_up)
    for i in range(idx_up, nb_layers):
        if i <= idx_down:
            tau[i] += (i + 1 - idx_up) * step
        else:
            tau[i] += tauStar_aer
    return tau


def mu_approx_In(mu, nb_angles):
    """SOS_Aer_I1_In.py:274-282."""
    i = nb_angles
    while mu[i] < 0.009:
        i += 1
    first = i
    while mu[i] < 0.020:
        i += 1
    return first, i


def extrapolation_width(tau_ref: float, nb_angles: int) -> int:
    """SOS_Aer_I1_In.py:124-127 / SOS_Aer_main_specular.py:342-345."""
    if tau_ref <= 0.0625:
        return int(0.005 * nb_angles)
    if tau_ref <= 1:
        return int(0.02 * nb_angles)
    if tau_ref < 4:
        return int(0.04 * nb_angles)
    return int(0.06 * nb_angles)


# --------------------------------------------------------------------------
# mu -> 0 helpers (SOS_Aer_In_limit.py:70-141 == SOS_Aer_I1_In.py:199-270)
# --------------------------------------------------------------------------
def asymptotic_down(J_col, tau_col, tau_t, mu):
    """|mu| < MU_THRESHOLD branch (SOS_Aer_In_limit.py:70-109)."""
    if len(tau_col) == 0:
        return 0.0
    if abs(mu) < MU_VERY_SMALL_THRESHOLD:  # the 1e-8 and 1e-3 branches are identical (:79-93)
        slope = 0.0
        if len(tau_col) > 1:
            slope = (J_col[-1] - J_col[-2]) / (tau_col[-1] - tau_col[-2])
        return -J_col[-1] + mu * slope
    keep = np.where(tau_col >= (tau_t - 5 * abs(mu)))[0]
    if len(keep) == 0:
        return -J_col[-1]
    with np.errstate(all="ignore"):
        f = J_col[keep] * np.exp((tau_t - tau_col[keep]) / mu)
    if np.any(np.isinf(f)) or np.any(np.isnan(f)):
        return -J_col[-1]
    return -trapz(f, tau_col[keep]) / mu


def limit_mu_down(row_down, mu_down, idx, i):
    """improved_limit_mu_down (SOS_Aer_In_limit.py:113-141)."""
    npts = min(5, idx)
    if npts < 2:
        slope = (row_down[-idx - 2] - row_down[-idx - 1]) / (mu_down[-idx - 2] - mu_down[-idx - 1])
        return slope * (mu_down[-i - 1] - mu_down[-idx - 1]) + row_down[-idx - 1]
    xs = mu_down[-(idx + npts):-idx]
    ys = row_down[-(idx + npts):-idx]
    if len(xs) >= 3:
        c = np.polyfit(np.array(xs, dtype=np.float64), np.array(ys, dtype=np.float64), 2)
        return np.polyval(c, float(mu_down[-i - 1]))
    slope = (ys[-1] - ys[0]) / (xs[-1] - xs[0])
    return ys[0] + slope * (mu_down[-i - 1] - xs[0])


def extrapolation_matrix(mu_down, idx):
    """The fixed linear map of limit_mu_down: target[i] = sum_j W[i,j]*source[j].

    Returns (W (idx, ns), src0, ns): sources are columns src0 .. src0+ns-1 of the
    downward half, target i is column M-1-i.  Built by pushing unit vectors through
    limit_mu_down so that np.polyfit's own scaling/lstsq is what defines W.
    """
    M = len(mu_down)
    if idx <= 0:
        return np.zeros((0, 0)), 0, 0
    npts = min(5, idx)
    if npts < 2:
        src0, ns = M - idx - 2, 2
    else:
        src0, ns = M - idx - npts, npts
    W = np.zeros((idx, ns))
    for j in range(ns):
        e = np.zeros(M)
        e[src0 + j] = 1.0
        for i in range(idx):
            W[i, j] = limit_mu_down(e, mu_down, idx, i)
    return W, src0, ns


# --------------------------------------------------------------------------
# single homogeneous layer: the literal drop-in functions
# --------------------------------------------------------------------------
def I1_NumInt(tau, mu, tauStar, mu0, P0, alb, nb_angles):
    """SOS_Aer_I1_In.py:13-58."""
    M = nb_angles
    tau = np.asarray(tau, dtype=np.float64)
    L = len(tau)
    out = np.zeros((L, 2 * M))
    e0 = np.exp(-tau / mu0)
    eS = np.exp(-tauStar / mu0)
    k = alb / (4 * np.pi)
    with np.errstate(all="ignore"):
        # downward general columns (:34-37)
        md = mu[: M - 1]
        out[:, : M - 1] = (mu0 / (mu0 + md)) * k * P0[: M - 1] * (e0[:, None] - np.exp(tau[:, None] / md[None, :]))
        # mu = 0- (:39)
        out[:, M - 1] = k * (mu0 / (mu0 + mu[M - 1])) * P0[M - 1] * e0
        # |mu + mu0| < 1e-4 (:41-43)
        hit = np.abs(md + mu0) < MU0_TOLERANCE
        if np.any(hit):
            out[:, : M - 1][:, hit] = k * P0[: M - 1][hit] * (e0 * tau / mu0)[:, None]
        # upward (:50-55)
        out[:, M] = k * (mu0 / (mu0 + mu[M])) * P0[M] * e0
        mp = mu[M + 1:]
        out[:, M + 1:] = (mu0 / (mu0 + mp)) * k * P0[M + 1:] * (
            e0[:, None] - eS * np.exp(-(tauStar - tau)[:, None] / mp[None, :]))
    return out * np.pi / mu0


def Jn_NumInt(n, In_1, tau, mu, tauStar, mu0, P, alb, nb_angles):
    """SOS_Aer_I1_In.py:62-74: J[t,m] = alb/4 * trapz_k(P[m, N-1-k] * I[t,k], mu)."""
    L = len(tau)
    out = np.zeros((L, 2 * nb_angles))
    Pf = P[:, ::-1]
    for t in range(L):
        out[t, :] = (alb / 4) * trapz(Pf * In_1[t, :], mu, axis=1)
    return out


def In_NumInt(n, Jn, In_1, tau, mu, tauStar, mu0, P, alb, nb_angles, mu_1=None, mu_2=None, method="slices"):
    """SOS_Aer_I1_In.py:77-130 (one homogeneous layer, no surface)."""
    tau = np.asarray(tau, dtype=np.float64)
    lay = Layout(tau=tau, mu=np.asarray(mu, dtype=np.float64), nb_angles=nb_angles,
                 regions=[(0, len(tau))], tau_ref=[float(tauStar)],
                 thick=bool(tauStar / mu[nb_angles + 1] >= 50), surface="none", grd_alb=0.0)
    return order_sweeps(lay, Jn, method=method)


# --------------------------------------------------------------------------
# general layered sweeps (1 region = single layer, 3 regions = the drivers)
# --------------------------------------------------------------------------
@dataclass
class Layout:
    tau: np.ndarray
    mu: np.ndarray
    nb_angles: int
    regions: List[Tuple[int, int]]  # [r0, r1) row ranges, top to bottom
    tau_ref: List[float]  # per region, selects the extrapolation width
    thick: bool
    surface: str  # 'none' | 'specular' | 'lambert'
    grd_alb: float


class BlendSearchOverrun(IndexError):
    """The reference's unbounded second-difference search ran off the row (Q11)."""


def _blend_row(row, mu, M):
    """SOS_Aer_I1_In.py:101-108: find-first on the raw row, then lerp towards mu=0+."""
    N = 2 * M
    i = M + 1
    while True:
        if i + 2 > N - 1:
            raise BlendSearchOverrun(f"blend search overran the row (N={N})")
        if not (abs((row[i] - row[i + 1]) - (row[i + 1] - row[i + 2])) > BLEND_THRESHOLD):
            break
        i += 1
    i += 1
    for m in range(M + 1, i):
        w = mu[m] / mu[i]
        row[m] = (1 - w) * row[M] + w * row[i]
    return i


def _extrapolate_row(row, mu, M, idx):
    for i in range(idx):
        row[M - 1 - i] = limit_mu_down(row[:M], mu[:M], idx, i)


def order_sweeps(lay: Layout, J: np.ndarray, method="slices", blend_index_out=None) -> np.ndarray:
    """One order's layer integration: J (L,N) -> I_n (L,N).

    Down sweep per region then surface coupling then up sweep per region from the
    bottom (SOS_Aer_main_specular.py:327-449; SOS_Aer_main_lambertian.py:399,401).
    """
    tau, mu, M = lay.tau, lay.mu, lay.nb_angles
    L, N = len(tau), 2 * M
    out = np.zeros((L, N))
    small = np.abs(mu[: M - 1]) < MU_THRESHOLD
    std = np.where(~small)[0]
    asy = np.where(small)[0]
    mstd = mu[std]

    # ---------------- downward ----------------
    if method == "recurrence":
        # standard columns: one continuous recurrence over all rows (region boundaries
        # are invisible because every region's slice starts at its carry row)
        raw = np.zeros(len(std))
        out[0, std] = raw
        for t in range(1, L):
            d = tau[t] - tau[t - 1]
            a = np.exp(d / mstd)
            raw = raw * a - (d / 2.0) * (J[t - 1, std] * a + J[t, std]) / mstd
            out[t, std] = raw
    for k, (r0, r1) in enumerate(lay.regions):
        idx = extrapolation_width(lay.tau_ref[k], M)
        c = r0 - 1 if k > 0 else 0  # first row of the slice (= carry row for k > 0)
        for t in range(r0, r1):
            if method == "slices":
                ts = tau[c: t + 1]
                E = np.exp((tau[t] - ts)[:, None] / mstd[None, :])
                integ = trapz(J[c: t + 1][:, std] * E, ts, axis=0)
                if k == 0:
                    out[t, std] = -integ / mstd
                else:
                    out[t, std] = out[c, std] * np.exp((tau[t] - tau[c]) / mstd) - integ / mstd
            for m in asy:
                out[t, m] = asymptotic_down(J[r0: t + 1, m], tau[r0: t + 1], tau[t], mu[m])
            _extrapolate_row(out[t], mu, M, idx)

    # ---------------- surface ----------------
    last = L - 1
    if lay.surface == "specular":
        seed = lay.grd_alb * out[last, M - 2::-1]  # mirror of cols M+1..N-1 is M-2..0
    elif lay.surface == "lambert":
        cols = np.arange(M - 2, -1, -1)
        seed = np.full(M - 1, -2 * lay.grd_alb * trapz(out[last, cols] * mu[cols], mu[cols]))
    elif lay.surface == "lambert_readme":
        # README.md:215 (not the shipped code): -2 rho int_{-1}^{0} I mu dmu, ascending abscissa, whole downward half
        seed = np.full(M - 1, -2 * lay.grd_alb * trapz(out[last, :M] * mu[:M], mu[:M]))
    else:
        seed = np.zeros(M - 1)

    # ---------------- upward ----------------
    up = np.arange(M + 1, N)
    mup = mu[up]
    R = len(lay.regions)
    for k in range(R - 1, -1, -1):
        r0, r1 = lay.regions[k]
        is_last = k == R - 1
        end = L if is_last else r1  # slice t:end
        b_row = last if is_last else r1
        if method == "recurrence":
            raw_next = None
            for t in range(r1 - 1, r0 - 1, -1):
                if is_last and t == last:
                    raw = seed.copy()  # zero-length integral
                elif t == end - 1:
                    # topmost row of the carry gap: pure attenuation (A.7, s_t = 0)
                    raw = out[b_row, up] * np.exp(-(tau[b_row] - tau[t]) / mup)
                else:
                    d = tau[t + 1] - tau[t]
                    a = np.exp(-d / mup)
                    raw = raw_next * a + (d / 2.0) * (J[t, up] + J[t + 1, up] * a) / mup
                raw_next = raw
                out[t, up] = raw
        for t in range(r0, r1):
            if method == "slices":
                ts = tau[t:end]
                E = np.exp(-(ts - tau[t])[:, None] / mup[None, :])
                bnd = (seed if is_last else out[b_row, up]) * np.exp(-(tau[b_row] - tau[t]) / mup)
                if lay.thick:
                    out[t, up] = bnd + trapz(J[t:end][:, up] * (E / mup), ts, axis=0)
                else:
                    out[t, up] = bnd + trapz(J[t:end][:, up] * E, ts, axis=0) / mup
        # blend after the raw rows of the region are known; the carry row of the next
        # region up is therefore read post-blend (A.7 "re-seeding")
        for t in range(r0, r1):
            out[t, M] = J[t, M]
            bi = _blend_row(out[t], mu, M)
            if blend_index_out is not None:
                blend_index_out[t] = bi
    return out


# NOTE on the recurrence method and blending: inside a region the reference's
# slice integral always uses raw J and the (blended) carry row only, so the raw
# recurrence must run on raw values; the loop above keeps `raw_next` separate
# from `out` for exactly that reason and blends only after the region is done.


# --------------------------------------------------------------------------
# three-region drivers
# --------------------------------------------------------------------------
@dataclass
class Scenario:
    """Parameters of SOS_Aer() (SOS_Aer_main_specular.py:23-94)."""
    mu0: float = 0.5
    z0: float = 120.0
    z_up: float = 25.0
    z_down: float = 17.0
    nb_layers: int = 800
    tauStar_atm: float = 0.104
    tauStar_aer: float = 0.120
    grd_alb: float = 1.0
    alb_atm: float = 1.0
    alb_aer: float = 1.0
    nb_angles: int = 501
    surface: str = "specular"  # 'specular' | 'lambert' (Lambert-as-coded, repair A)
    threshold: float = CONVERGENCE
    max_orders: int = 10000

    def geometry(self):
        z_up, z_down = self.z_up, self.z_down
        if z_down > z_up:
            z_down, z_up = z_up, z_down
        L = self.nb_layers
        tau = tau_profile(self.tauStar_atm, self.tauStar_aer, self.z0, z_up, z_down, L)
        z = np.linspace(self.z0, 0, L)
        idx_up = int(np.argmin(np.abs(z - z_up)))
        idx_down = int(np.argmin(np.abs(z - z_down)))
        return tau, z, idx_up, idx_down


def first_order_regions(sc: Scenario, tau, mu, idx_up, idx_down, P0_atm, P0_aer):
    """Closed-form first order over 3 regions (SOS_Aer_main_specular.py:104-292)."""
    M, L = sc.nb_angles, sc.nb_layers
    N = 2 * M
    mu0 = sc.mu0
    F0 = np.pi / mu0
    Tstar = sc.tauStar_atm + sc.tauStar_aer
    dtau_aer = sc.tauStar_aer / (idx_down + 1 - idx_up)
    dtau_atm = sc.tauStar_atm / L
    f_atm = dtau_atm / (dtau_atm + dtau_aer)
    f_aer = dtau_aer / (dtau_atm + dtau_aer)
    C_atm = sc.alb_atm * P0_atm
    C_mix = sc.alb_atm * P0_atm * f_atm + sc.alb_aer * P0_aer * f_aer
    S = F0 * sc.grd_alb * np.exp(-Tstar / mu0)
    q = 1.0 / (4 * np.pi)
    I1 = np.zeros((L, N))
    mir = lambda m: N - 1 - m
    regions = [(0, idx_up, C_atm), (idx_up, idx_down + 1, C_mix), (idx_down + 1, L, C_atm)]

    with np.errstate(all="ignore"):
        # ---- downward, top region first ----
        md = np.arange(M - 1)
        mud = mu[md]
        hit = np.abs(mud + mu0) < MU0_TOLERANCE
        for k, (r0, r1, C) in enumerate(regions):
            for t in range(r0, r1):
                if k == 0:
                    carry, tau_d, tau_s = 0.0, 0.0, 0.0
                else:
                    c = r0 - 1
                    carry = I1[c, md] * np.exp((tau[t] - tau[c]) / mud)
                    tau_d, tau_s = tau[c], tau[r0]
                direct = (mu0 / (mu0 + mud)) * C[md] * (F0 * q) * (
                    np.exp(-tau[t] / mu0) - np.exp(-tau_d / mu0) * np.exp((tau[t] - tau_d) / mud))
                surf = (mu0 / (mu0 - mud)) * C[mir(md)] * (S * q) * (
                    np.exp(-(Tstar - tau[t]) / mu0) - np.exp(-(Tstar - tau_s) / mu0) * np.exp((tau[t] - tau_s) / mud))
                row = carry + direct + surf
                if np.any(hit):
                    d2 = C[md[hit]] * (F0 * q) * np.exp(-tau[t] / mu0) * (tau[t] - tau_d) / mu0
                    cc = carry[hit] if k > 0 else 0.0
                    row[hit] = cc + d2 + surf[hit]
                I1[t, md] = row
                I1[t, M - 1] = (mu0 / (mu0 + mu[M - 1])) * C[M - 1] * (F0 * q) * np.exp(-tau[t] / mu0) \
                    + (mu0 / (mu0 - mu[M - 1])) * C[M] * (S * q) * np.exp(-(Tstar - tau[t]) / mu0)
        # ---- upward, bottom region first ----
        mu_ = np.arange(M + 1, N)
        muu = mu[mu_]
        hit = np.abs(muu - mu0) < MU0_TOLERANCE
        for k in (2, 1, 0):
            r0, r1, C = regions[k]
            for t in range(r0, r1):
                if k == 2:
                    b = L - 1
                    carry = sc.grd_alb * I1[b, mir(mu_)] * np.exp(-(tau[b] - tau[t]) / muu)
                    tau_d, tau_s = tau[b], Tstar
                else:
                    b = r1
                    carry = I1[b, mu_] * np.exp(-(tau[b] - tau[t]) / muu)
                    tau_d, tau_s = tau[b], tau[r1 - 1]
                direct = (mu0 / (mu0 + muu)) * C[mu_] * (F0 * q) * (
                    np.exp(-tau[t] / mu0) - np.exp(-tau_d / mu0) * np.exp(-(tau_d - tau[t]) / muu))
                surf = (mu0 / (mu0 - muu)) * C[mir(mu_)] * (S * q) * (
                    np.exp(-(Tstar - tau[t]) / mu0) - np.exp(-(Tstar - tau_s) / mu0) * np.exp(-(tau_s - tau[t]) / muu))
                row = carry + direct + surf
                if np.any(hit):
                    s2 = C[mir(mu_[hit])] * (S * q) * np.exp(-(Tstar - tau[t]) / mu0) * (tau_s - tau[t]) / mu0
                    row[hit] = carry[hit] + direct[hit] + s2
                I1[t, mu_] = row
                I1[t, M] = (mu0 / (mu0 + mu[M])) * C[M] * (F0 * q) * np.exp(-tau[t] / mu0) \
                    + (mu0 / (mu0 - mu[M])) * C[M - 1] * (S * q) * np.exp(-(Tstar - tau[t]) / mu0)
    return I1


def source_regions(sc: Scenario, In_1, mu, idx_up, idx_down, P_atm, P_aer):
    """SOS_Aer_main_specular.py:315-323."""
    L = sc.nb_layers
    dtau_aer = sc.tauStar_aer / (idx_down + 1 - idx_up)
    dtau_atm = sc.tauStar_atm / L
    f_atm = dtau_atm / (dtau_atm + dtau_aer)
    f_aer = dtau_aer / (dtau_atm + dtau_aer)
    J = np.zeros_like(In_1)
    Pa, Pe = P_atm[:, ::-1], P_aer[:, ::-1]
    for t in range(L):
        ja = (sc.alb_atm / 4) * trapz(Pa * In_1[t, :], mu, axis=1)
        if idx_up <= t <= idx_down:
            J[t] = ja * f_atm + (sc.alb_aer / 4) * trapz(Pe * In_1[t, :], mu, axis=1) * f_aer
        else:
            J[t] = ja
    return J


def source_gemm(In_1, A):
    """The same contraction as one matrix product."""
    return In_1 @ A


def contraction_matrix(P, mu, alb):
    """A[k,m] = (alb/4) * w_k * P[m, N-1-k], w = composite-trapezoid weights on mu."""
    d = np.diff(mu)
    w = np.zeros_like(mu)
    w[:-1] += d / 2
    w[1:] += d / 2
    return (alb / 4) * (w[:, None] * P[:, ::-1].T)


def driver_layout(sc: Scenario, tau, mu, idx_up, idx_down) -> Layout:
    L, M = sc.nb_layers, sc.nb_angles
    return Layout(
        tau=tau, mu=mu, nb_angles=M,
        regions=[(0, idx_up), (idx_up, idx_down + 1), (idx_down + 1, L)],
        tau_ref=[float(tau[idx_up - 1]), float(tau[idx_down]), float(tau[idx_down])],
        thick=bool(tau[L - 1] / mu[M + 1] >= 50),
        surface=sc.surface, grd_alb=sc.grd_alb)


def convergence_ratio(In, I, M):
    """SOS_Aer_main_specular.py:309 (Python max over NumPy scalars)."""
    L = I.shape[0]
    with np.errstate(all="ignore"):
        return max(max(In[0, M:] / I[0, M:]), max(In[L - 1, :M] / I[L - 1, :M]))


def solve(sc: Scenario, P0_atm, P_atm, P0_aer, P_aer, method="slices", keep_orders=True, use_gemm=False):
    """The whole SOS_Aer() solve; returns dict(I, I_saved, n, tau, mu, z, idx_up, idx_down, ratios)."""
    tau, z, idx_up, idx_down = sc.geometry()
    M = sc.nb_angles
    mu = mu_grid(M)
    I1 = first_order_regions(sc, tau, mu, idx_up, idx_down, P0_atm, P0_aer)
    lay = driver_layout(sc, tau, mu, idx_up, idx_down)
    if use_gemm:
        L = sc.nb_layers
        dtau_aer = sc.tauStar_aer / (idx_down + 1 - idx_up)
        dtau_atm = sc.tauStar_atm / L
        f_atm = dtau_atm / (dtau_atm + dtau_aer)
        f_aer = dtau_aer / (dtau_atm + dtau_aer)
        A_atm = contraction_matrix(P_atm, mu, sc.alb_atm)
        A_mix = f_atm * A_atm + f_aer * contraction_matrix(P_aer, mu, sc.alb_aer)
    In_1 = I1
    I = I1.copy()
    saved = [I1]
    In = np.ones_like(I1)
    n = 1
    ratios = []
    while True:
        r = convergence_ratio(In, I, M)
        if not (r >= sc.threshold) or n >= sc.max_orders:
            break
        ratios.append(float(r))
        n += 1
        if use_gemm:
            J = In_1 @ A_atm
            J[idx_up: idx_down + 1] = In_1[idx_up: idx_down + 1] @ A_mix
        else:
            J = source_regions(sc, In_1, mu, idx_up, idx_down, P_atm, P_aer)
        In = order_sweeps(lay, J, method=method)
        In_1 = In
        I = I + In
        if keep_orders:
            saved.append(In)
    return dict(I=I, I_saved=saved, n=n, tau=tau, mu=mu, z=z, idx_up=idx_up, idx_down=idx_down,
                ratios=ratios, final_ratio=float(r))


# --------------------------------------------------------------------------
# quadratures (SOS_Aer_graphe.py)
# --------------------------------------------------------------------------
def flux_up_down(I, mu, M, tau, mu0, F0, grd_alb, direct_scale=1.0):
    """SOS_Aer_graphe.py:154-158 (direct_scale=1) and :74-78 / SOS_Aer_critical_albedo.py:377-381
    (direct_scale=1/(4 pi))."""
    L = I.shape[0]
    down = trapz(I[:, :M] * mu[:M], mu[:M], axis=1) - (F0 * direct_scale) * np.exp(-tau / mu0)
    up = trapz(I[:, M:] * mu[M:], mu[M:], axis=1) + (F0 * direct_scale) * grd_alb * np.exp(-(2 * tau[L - 1] - tau) / mu0)
    return up, down


def net_flux(I, mu, tau, mu0, F0, grd_alb):
    """SOS_Aer_graphe.py:39-41."""
    L = I.shape[0]
    return trapz(I * mu, mu, axis=1) - F0 * np.exp(-tau / mu0) + grd_alb * F0 * np.exp(-(2 * tau[L - 1] - tau) / mu0)


def diffusivity(I, mu):
    """SOS_Aer_graphe.py:8-10."""
    with np.errstate(all="ignore"):
        return -trapz(I * mu, mu, axis=1) / trapz(I, mu, axis=1)


def heating_rate(I, mu, z, M, idx_up, idx_down, F0, mu0, tau, grd_alb):
    """SOS_Aer_graphe.py:70-91."""
    rho, cp = 1.225, 1004
    up, down = flux_up_down(I, mu, M, tau, mu0, F0, grd_alb, direct_scale=1.0 / (4 * np.pi))
    f = down + up
    L = I.shape[0]
    hr = np.zeros(L)
    hr[:-1] = -(1 / (rho * cp)) * (f[1:] - f[:-1]) / (z[1:] - z[:-1])
    hr[-1] = hr[-2]
    hr[idx_up - 1] = hr[idx_up - 2]
    hr[idx_down] = hr[idx_down - 1]
    return hr


def toa_net_flux(I, mu, M, tau, mu0, F0, grd_alb):
    """SOS_Aer_critical_albedo.py:377-382."""
    up, down = flux_up_down(I, mu, M, tau, mu0, F0, grd_alb, direct_scale=1.0 / (4 * np.pi))
    return -down[0] - up[0]
