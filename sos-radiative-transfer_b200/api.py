"""Drop-in single-layer operators with the reference's exact positional signatures.

    I1_NumInt(tau, mu, tauStar, mu0, P0, alb, nb_angles)                       SOS_Aer_I1_In.py:13
    Jn_NumInt(n, In_1, tau, mu, tauStar, mu0, P, alb, nb_angles)               SOS_Aer_I1_In.py:62
    In_NumInt(n, Jn, In_1, tau, mu, tauStar, mu0, P, alb, nb_angles, mu_1, mu_2)   SOS_Aer_I1_In.py:77
    mu_approx_In(mu, nb_angles)                                                SOS_Aer_I1_In.py:274

NumPy arrays in, a fresh C-contiguous float64 (len(tau), 2*nb_angles) array out, inputs never
modified.  The arguments the reference ignores (n, tauStar/mu0 in Jn_NumInt; n, In_1, mu0, P, alb,
mu_1, mu_2 in In_NumInt) are accepted and ignored here as well.  Device tensors (torch, float64,
CUDA) are accepted for In_1 / Jn and then returned as device tensors without a host round trip.
The reference's IndexError from the unbounded blend search (SOS_Aer_I1_In.py:103) is raised as
IndexError too.
"""
from __future__ import annotations

import hashlib
from collections import OrderedDict

import numpy as np
import torch

from . import _lib
from . import grid as G
from .engine import ScenarioCoefficients, SosEngine

_CACHE: "OrderedDict[tuple, SosEngine]" = OrderedDict()
_CACHE_MAX = 4


def _digest(a: np.ndarray) -> bytes:
    return hashlib.blake2b(np.ascontiguousarray(a).view(np.uint8), digest_size=16).digest()


def _engine(tau, mu, tauStar, mu0, alb, nb_angles, P=None) -> SosEngine:
    _lib.load()
    if not torch.cuda.is_available():
        raise _lib.SosError("no CUDA device: the SOS engine has no CPU fallback")
    tau = np.ascontiguousarray(tau, dtype=np.float64)
    mu = np.ascontiguousarray(mu, dtype=np.float64)
    key = (int(nb_angles), float(tauStar), float(mu0), float(alb), _digest(tau), _digest(mu),
           torch.cuda.current_device())
    eng = _CACHE.get(key)
    if eng is None:
        w = G.extrapolation_width(float(tauStar), int(nb_angles))
        sc = ScenarioCoefficients(mu0=float(mu0), grd_alb=0.0, tauStar_tot=float(tauStar), coef_atm=float(alb),
                                  extrap_width=(w, w, w))
        eng = SosEngine(mu, tau[None, :], [sc], region_start=[0, len(tau)], surface=_lib.SURFACE_NONE)
        eng._phase_digest = None
        _CACHE[key] = eng
        while len(_CACHE) > _CACHE_MAX:
            _CACHE.popitem(last=False)[1].close()
    _CACHE.move_to_end(key)
    if P is not None:
        dg = _digest(np.asarray(P, dtype=np.float64))
        if eng._phase_digest != dg:
            eng.set_phase([P])
            eng._phase_digest = dg
    return eng


def clear_cache():
    while _CACHE:
        _CACHE.popitem()[1].close()


def mu_approx_In(mu, nb_angles):
    return G.mu_approx_In(mu, nb_angles)


def I1_NumInt(tau, mu, tauStar, mu0, P0, alb, nb_angles):
    eng = _engine(tau, mu, tauStar, mu0, alb, nb_angles)
    Cc = np.zeros((1, 2, eng.N))
    Cc[0, 0] = alb * np.asarray(P0, dtype=np.float64)
    return eng.to_host(eng.first_order(Cc))


def Jn_NumInt(n, In_1, tau, mu, tauStar, mu0, P, alb, nb_angles):
    eng = _engine(tau, mu, tauStar, mu0, alb, nb_angles, P=P)
    on_device = isinstance(In_1, torch.Tensor) and In_1.is_cuda
    J = eng.source(eng.to_field(In_1))
    return J if on_device else eng.to_host(J)


def In_NumInt(n, Jn, In_1, tau, mu, tauStar, mu0, P, alb, nb_angles, mu_1=None, mu_2=None):
    eng = _engine(tau, mu, tauStar, mu0, alb, nb_angles)
    on_device = isinstance(Jn, torch.Tensor) and Jn.is_cuda
    eng.reset(eng.to_field(np.ones((eng.L, eng.N))))  # all scenarios active, status cleared
    out = eng.sweeps(eng.to_field(Jn))
    res = eng.results()
    if res[0].status & _lib.STATUS_BLEND_OVERRUN:
        raise IndexError("In_NumInt: mu->0 blend search ran off the row "
                         "(the reference raises IndexError at SOS_Aer_I1_In.py:103)")
    return out if on_device else eng.to_host(out)
