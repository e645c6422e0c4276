"""The reference's output layer (SOS_Aer_graphe.py) with its own call signatures, computed on the device.

    graphe_diffusivity(I, mu, z_profile, nb_layers, aer_phase_fun)                                     :6
    graphe_flux(I, mu, z_profile, nb_layers, nb_angles, tau, mu0, F0, grd_alb, aer_phase_fun)          :37
    graphe_heating_rate(I, mu, z_profile, nb_layers, nb_angles, idx_up, idx_down, F0, mu0, tau, grd_alb, aer_phase_fun)  :68
    graphe_successive_dif(I_saved, mu, z_profile, nb_layers, nb_angles, aer_phase_fun)                 :118
    graphe_flux_up_down(I, mu, z_profile, nb_layers, nb_angles, tau, mu0, F0, grd_alb, aer_phase_fun)  :152

The reference functions compute one curve per call with a Python loop over the layers (two np.trapz per layer), plot it
and save a PNG into a hard-coded Windows folder; they return nothing.  Here the quadratures run on the GPU
(csrc/quadrature.cuh through sos_quadratures: the same kernels the drivers use) and every function RETURNS the curve(s)
it would have plotted.  Plotting is optional: plot=True draws the same figure with matplotlib when it is installed
(no file is written unless save=<path>); the hot path never needs it.

successive_diffusivity(I_saved, mu) is the batched form: all saved orders are stacked as one batch and reduced in a
single launch (the per-order mean diffusivity of SOS_Aer_graphe.py:118-149).
"""
from __future__ import annotations

from typing import Optional, Sequence

import numpy as np
import torch

from . import _lib
from . import grid as G
from .engine import ScenarioCoefficients, SosEngine


def _engine(mu, tau, mu0, grd_alb, n_fields, regions=None) -> SosEngine:
    """A plan that only serves the quadrature kernels: n_fields copies of the same (tau, mu0, albedo)."""
    _lib.load()
    if not torch.cuda.is_available():
        raise _lib.SosError("no CUDA device: the SOS engine has no CPU fallback")
    mu = np.ascontiguousarray(mu, dtype=np.float64)
    tau = np.ascontiguousarray(tau, dtype=np.float64)
    L = tau.shape[-1]
    w = G.extrapolation_width(float(tau[-1]), len(mu) // 2)   # (a plan wants valid widths; the quadratures ignore them)
    coefs = [ScenarioCoefficients(mu0=float(mu0), grd_alb=float(grd_alb), tauStar_tot=float(tau[-1]), coef_atm=1.0,
                                  extrap_width=(w, w, w)) for _ in range(n_fields)]
    region_start = [0, L] if regions is None else [0, int(regions[0]), int(regions[1]) + 1, L]
    return SosEngine(mu, np.tile(tau, (n_fields, 1)), coefs, region_start, _lib.SURFACE_NONE)


def _curves(fields: Sequence[np.ndarray], mu, tau, mu0, F0, grd_alb, z=None, regions=None):
    """All quadratures of a stack of (L, N) fields sharing one tau profile: dict of (n_fields, L) arrays."""
    fields = [np.asarray(f, dtype=np.float64) for f in fields]
    L, N = fields[0].shape
    for f in fields:
        if f.shape != (L, N):
            raise ValueError("Error in I_saved memory, not the good shape")   # (the reference prints this, :126)
    eng = _engine(mu, tau, mu0, grd_alb, len(fields), regions)
    try:
        dev = eng.to_field(np.stack(fields))
        # the kernels use F0 = pi / mu0 (SOS_Aer_main_specular.py:30); another F0 rescales the direct terms of the fluxes
        scale = float(F0) / (np.pi / float(mu0))
        q = eng.quadratures(dev, z, direct_scale=scale, heating=z is not None and regions is not None)
    finally:
        eng.close()
    return q


def _maybe_plot(xs, z_profile, labels, xlabel, title, plot, save):
    if not plot and not save:
        return
    import matplotlib.pyplot as plt   # optional dependency: only the plotting shim needs it
    for x, lab in zip(xs, labels):
        plt.plot(x, z_profile, label=lab)
    plt.xlabel(xlabel)
    plt.ylabel("Altitude (km)")
    plt.title(title)
    plt.grid(True)
    if any(labels):
        plt.legend()
    if save:
        plt.savefig(save, dpi=600)
    if plot:
        plt.show()
    plt.close()


def successive_diffusivity(I_saved: Sequence[np.ndarray], mu) -> np.ndarray:
    """(n_orders, L): -trapz(I_n mu, mu) / trapz(I_n, mu) for every saved order (SOS_Aer_graphe.py:129-131), one launch."""
    L = np.asarray(I_saved[0]).shape[0]
    q = _curves(I_saved, mu, np.zeros(L), 1.0, np.pi, 0.0)
    return q["diffusivity"]


def graphe_successive_dif(I_saved, mu, z_profile, nb_layers, nb_angles, aer_phase_fun, plot=False, save=None):
    for I in I_saved:
        if np.asarray(I).shape != (nb_layers, 2 * nb_angles):
            raise ValueError("Error in I_saved memory, not the good shape")
    dif = successive_diffusivity(I_saved, mu)
    _maybe_plot(dif, z_profile, [f"order={m + 1}" for m in range(len(dif))], r"Diffusivity $\bar{\mu}$",
                rf"Diffusivity $\bar{{\mu}}$ for {aer_phase_fun} layer (SOS)", plot, save)
    return dif


def graphe_diffusivity(I, mu, z_profile, nb_layers, aer_phase_fun, plot=False, save=None):
    """(L,): -trapz(I mu, mu) / trapz(I, mu) per layer (SOS_Aer_graphe.py:8-10)."""
    dif = _curves([I], mu, np.zeros(nb_layers), 1.0, np.pi, 0.0)["diffusivity"][0]
    _maybe_plot([dif], z_profile, [""], r"Diffusivity $\bar{\mu}$", rf"Diffusivity $\bar{{\mu}}$ for {aer_phase_fun} layer", plot, save)
    return dif


def graphe_flux(I, mu, z_profile, nb_layers, nb_angles, tau, mu0, F0, grd_alb, aer_phase_fun, plot=False, save=None):
    """(L,): net flux trapz(I mu, mu) - F0 exp(-tau/mu0) + grd_alb F0 exp(-(2 tau* - tau)/mu0) (SOS_Aer_graphe.py:39-41)."""
    q = _curves([I], mu, tau, mu0, F0, grd_alb)
    flux = q["flux_down"][0] + q["flux_up"][0]   # (= the kernel's net flux when F0 = pi / mu0)
    _maybe_plot([flux], z_profile, [""], "Net flux", f"Net flux for {aer_phase_fun} layer", plot, save)
    return flux


def graphe_flux_up_down(I, mu, z_profile, nb_layers, nb_angles, tau, mu0, F0, grd_alb, aer_phase_fun, plot=False, save=None):
    """(flux_up, flux_down), each (L,) (SOS_Aer_graphe.py:154-158)."""
    q = _curves([I], mu, tau, mu0, F0, grd_alb)
    up, down = q["flux_up"][0], q["flux_down"][0]
    _maybe_plot([up, down], z_profile, ["flux up", "flux down"], "Flux", f"Upward and downward flux for {aer_phase_fun} layer", plot, save)
    return up, down


def graphe_heating_rate(I, mu, z_profile, nb_layers, nb_angles, idx_up, idx_down, F0, mu0, tau, grd_alb, aer_phase_fun,
                        plot=False, save=None):
    """(L,): heating rate -(1 / (rho cp)) dF/dz with the F0 / (4 pi) direct terms and the reference's three patched rows
    (SOS_Aer_graphe.py:70-91)."""
    if abs(float(F0) - np.pi / float(mu0)) > 1e-12 * abs(float(F0)):
        raise ValueError("graphe_heating_rate: the device kernel assumes F0 = pi / mu0 (SOS_Aer_main_specular.py:30)")
    q = _curves([I], mu, tau, mu0, F0, grd_alb, z=np.asarray(z_profile, dtype=np.float64), regions=(idx_up, idx_down))
    hr = q["heating_rate"][0]
    _maybe_plot([hr], z_profile, [""], "Heating rate (K/s)", f"Heating rate for {aer_phase_fun} layer", plot, save)
    return hr
