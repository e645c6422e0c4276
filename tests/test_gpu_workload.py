"""GPU parity of what bench.py actually times, and of the order loop with and without the generated source.

* the benchmarked workload itself (bench.make_scenarios: Rayleigh atmosphere + HG / log-normal-Mie / FWC aerosols in
  the three-region specular driver at 800 x 1002, mu0 in [0.1, 1], omega_aer in [0.7, 1]) against the oracle;
* the FWC cloud as the aerosol of the three-region driver against a fixture produced by the unmodified reference;
* sos_solve with the generated source (csrc/sweep.cuh SrcGen: J rebuilt from two projections per row on the molecular rows,
  I_n kept only where something reads it, two-column apply pass on odd grids) against the same solve with every J row
  written by the contraction kernels and read back (SOS_B200_GENSRC=0), on the same batches: all surfaces, ragged sizes,
  even and odd M, Taylor / windowed columns, single solves.

Tolerance 1e-10 relative (BASELINE.json north_star), same order counts.
"""
import ast
import dataclasses

import numpy as np
import pytest

from conftest import relelem, relmax

pytestmark = pytest.mark.gpu

TOL = 1e-10


@pytest.fixture(scope="module")
def sos():
    import torch
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    import sos_b200
    sos_b200._lib.load()
    return sos_b200


@pytest.fixture(scope="module")
def so():
    import sos_oracle
    return sos_oracle


def _oracle(so, sos, sc):
    M = sc.nb_angles
    mu = so.mu_grid(M)
    P0a, Pa = sos.phase_matrices(sc.atm_phase[0], M, mu, sc.mu0, sc.atm_phase[1])
    P0e, Pe = sos.phase_matrices(sc.aer_phase[0], M, mu, sc.mu0, sc.aer_phase[1])
    osc = so.Scenario(mu0=sc.mu0, z0=sc.z0, z_up=sc.z_up, z_down=sc.z_down, nb_layers=sc.nb_layers,
                      tauStar_atm=sc.tauStar_atm, tauStar_aer=sc.tauStar_aer, grd_alb=sc.grd_alb, alb_atm=sc.alb_atm,
                      alb_aer=sc.alb_aer, nb_angles=M, surface=sc.surface, threshold=sc.threshold)
    return so.solve(osc, P0a, Pa, P0e, Pe, method="recurrence", use_gemm=True, keep_orders=False), mu


def stratified_subset(scen):
    """Twelve workload members: every aerosol family at mu0 = 0.1 and 1.0 and at omega_aer = 0.7 and 1.0."""
    sel = []
    for fam in ("hg", "mie_lognormal", "fwc"):
        fs = [i for i, s in enumerate(scen) if s.aer_phase[0] == fam]
        for cond in (lambda s: abs(s.mu0 - 0.1) < 1e-9, lambda s: abs(s.mu0 - 1.0) < 1e-9,
                     lambda s: abs(s.alb_aer - 0.7) < 1e-9, lambda s: abs(s.alb_aer - 1.0) < 1e-9):
            c = [i for i in fs if cond(scen[i]) and i not in sel]
            assert c, "the workload no longer covers the corners of its sweep"
            sel.append(c[0])
    return sel


def test_benchmarked_workload_vs_oracle(sos, so, monkeypatch):
    """bench.py's 96-scenario batch, solved exactly as bench.py solves it (one BatchSolver, generated source), against
    the oracle on a stratified dozen of its members: radiances, order counts, fluxes, diffusivity, heating rate, TOA
    net flux.  Then the whole batch again with every source row stored: same orders, same fields."""
    import bench
    scen = bench.make_scenarios(sos, 96)
    bs = sos.BatchSolver(scen)
    assert bs.engine.gensrc_enabled, "the benchmarked batch must run with the generated source"
    res = bs.solve(poll_every=2)
    assert not np.any(res.status), res.status
    out = bs.results(res, quadratures=True, fields=True)
    bs.engine.close()
    assert 7 <= min(o.n for o in out) and max(o.n for o in out) <= 30
    for i in stratified_subset(scen):
        sc = scen[i]
        ref, mu = _oracle(so, sos, sc)
        M = sc.nb_angles
        got = out[i]
        assert got.n == ref["n"], (i, got.n, ref["n"])
        assert relmax(got.I, ref["I"]) < TOL, i
        assert relelem(got.I, ref["I"]) < 1e-8, i
        F0 = np.pi / sc.mu0
        up, down = so.flux_up_down(ref["I"], mu, M, ref["tau"], sc.mu0, F0, sc.grd_alb)
        assert relmax(got.flux_up, up) < TOL and relmax(got.flux_down, down) < TOL, i
        assert relmax(got.net_flux, so.net_flux(ref["I"], mu, ref["tau"], sc.mu0, F0, sc.grd_alb)) < TOL, i
        assert relmax(got.diffusivity, so.diffusivity(ref["I"], mu)) < TOL, i
        hr = so.heating_rate(ref["I"], mu, ref["z"], M, ref["idx_up"], ref["idx_down"], F0, sc.mu0, ref["tau"], sc.grd_alb)
        scale = np.max(np.abs(down)) / abs(ref["z"][1] - ref["z"][0]) / (1.225 * 1004)
        assert np.max(np.abs(got.heating_rate - hr)) < TOL * scale, i
        toa = so.toa_net_flux(ref["I"], mu, M, ref["tau"], sc.mu0, F0, sc.grd_alb)
        assert abs(got.toa_net_flux - toa) < TOL * abs(toa), i
    # the same batch with every source row written and read back must tell the same story for all 96
    monkeypatch.setenv("SOS_B200_GENSRC", "0")
    bs2 = sos.BatchSolver(scen)
    assert not bs2.engine.gensrc_enabled
    res2 = bs2.solve(poll_every=2)
    out2 = bs2.results(res2, quadratures=False, fields=True)
    bs2.engine.close()
    for a, b in zip(out, out2):
        assert a.n == b.n
        assert relmax(a.I, b.I) < 1e-12


@pytest.mark.parametrize("tag", ["mu01", "mu1", "mid"])
def test_fwc_aerosol_three_region_vs_golden(sos, golden, tag):
    d = golden("drivers_fwc3.npz")
    kw = ast.literal_eval(str(d[tag + "_kw"]))
    n_ref = int(d[tag + "_n"])
    r = sos.SOS_Aer_main_specular(keep_orders=n_ref, atm_phase=("rayleigh", 0.0), aer_phase=("fwc", 0.0), **kw)
    assert r.n == n_ref
    rows = d[tag + "_rows"]
    assert relmax(r.I[rows], d[tag + "_I_rows"]) < TOL
    assert relmax(r.I[::5], d[tag + "_I_sub"]) < TOL
    for j in range(n_ref):
        assert relmax(r.I_saved[j][rows], d[tag + "_order_rows"][j]) < TOL, j
    assert relmax(r.flux_up, d[tag + "_flux_up"]) < TOL and relmax(r.flux_down, d[tag + "_flux_down"]) < TOL
    assert relmax(r.net_flux, d[tag + "_net_flux"]) < TOL and relmax(r.diffusivity, d[tag + "_diffusivity"]) < TOL


def _solve_both(sos, scs, monkeypatch, keep=0, **kw):
    outs = {}
    for flag in ("1", "0"):
        monkeypatch.setenv("SOS_B200_GENSRC", flag)
        bs = sos.BatchSolver(scs, **kw)
        assert bs.engine.gensrc_enabled == (flag == "1"), flag
        res = bs.solve(keep_orders=keep)
        outs[flag] = bs.results(res, quadratures=True, keep_orders=keep)
        bs.engine.close()
    return outs["1"], outs["0"]


@pytest.mark.parametrize("surface", ["specular", "lambert"])
@pytest.mark.parametrize("L,M", [(96, 251), (130, 501), (77, 100), (60, 128), (64, 1201), (203, 300)])
def test_generated_source_equals_stored_source(sos, so, monkeypatch, surface, L, M):
    """Same batches with and without the generated source: ragged L and M (odd M: two-column apply pass, even M: the
    scalar one; partial column blocks), Taylor + windowed columns (M = 1201), all extrapolation widths, dense and rebuilt
    source rows, every order compared; one member also against the oracle."""
    aer = (("hg", 0.5), ("fwc", 0.0), ("hg", 0.8))
    scs = [sos.Scenario(nb_layers=L, nb_angles=M, mu0=(0.5, 0.23, 0.9, 1.0)[i % 4], tauStar_atm=(0.124, 0.05, 0.6, 0.3)[i % 4],
                        tauStar_aer=(0.12, 0.0, 0.9, 1.2)[i % 4], alb_aer=(0.97, 1.0, 0.8, 0.85)[i % 4], alb_atm=(1.0, 1.0, 0.9, 1.0)[i % 4],
                        grd_alb=(0.15, 0.0, 0.3, 0.5)[i % 4], atm_phase=("rayleigh", 0.0), aer_phase=aer[i % 3], surface=surface)
           for i in range(7)]
    keep = 4
    a, b = _solve_both(sos, scs, monkeypatch, keep=keep)
    for i, (x, y) in enumerate(zip(a, b)):
        assert x.n == y.n, (i, x.n, y.n)
        assert relmax(x.I, y.I) < 1e-12, i
        for j in range(min(keep + 1, x.n)):
            assert relmax(x.I_saved[j], y.I_saved[j]) < 1e-12, (i, j)
        assert relmax(x.flux_up, y.flux_up) < 1e-12 and relmax(x.heating_rate, y.heating_rate) < 1e-9
    ref, _ = _oracle(so, sos, scs[2])
    assert a[2].n == ref["n"] and relmax(a[2].I, ref["I"]) < TOL


def test_generated_source_single_solve_default_grid(sos, so, monkeypatch):
    """One scenario alone (the reference's own use: configs 1-3) runs with the generated source too: EVA specular on the
    default 800 x 1002 grid against the stored-source solve and the oracle."""
    sc = sos.Scenario(nb_layers=800, nb_angles=501, mu0=0.5, tauStar_atm=0.124, tauStar_aer=0.12, alb_aer=0.97, grd_alb=0.15,
                      atm_phase=("rayleigh", 0.0), aer_phase=("hg", 0.5), surface="specular")
    a, b = _solve_both(sos, [sc], monkeypatch)
    assert a[0].n == b[0].n and relmax(a[0].I, b[0].I) < 1e-12
    ref, _ = _oracle(so, sos, sc)
    assert a[0].n == ref["n"] and relmax(a[0].I, ref["I"]) < TOL


def test_generated_source_dense_atmosphere_and_single_layer(sos, so, monkeypatch):
    """No low-rank operand at all (HG atmosphere: every J row is read, 32 B per element), and the single-layer grid
    (one region, no surface) with a Rayleigh operand (every row generated, no dense tile at all)."""
    scs = [sos.Scenario(nb_layers=120, nb_angles=251, mu0=0.4 + 0.1 * i, tauStar_atm=0.2, tauStar_aer=0.1 * i, alb_aer=0.9, grd_alb=0.2,
                        atm_phase=("hg", 0.3), aer_phase=("hg", 0.7), surface="specular") for i in range(6)]
    a, b = _solve_both(sos, scs, monkeypatch)
    for x, y in zip(a, b):
        assert x.n == y.n and relmax(x.I, y.I) < 1e-12
    # single layer, S = 12 copies with different tau*: SosEngine level
    import torch
    L, M = 150, 200
    N = 2 * M
    mu = sos.mu_grid(M)
    S = 12
    ts = np.linspace(0.3, 6.0, S)
    tau = np.stack([np.linspace(0, t, L) for t in ts])
    out = {}
    for flag in ("1", "0"):
        monkeypatch.setenv("SOS_B200_GENSRC", flag)
        P0, P = sos.phase_matrices("rayleigh", M, mu, 0.6, 0.0)
        coefs = [sos.ScenarioCoefficients(mu0=0.6, grd_alb=0.0, tauStar_tot=float(t), coef_atm=0.95,
                                          extrap_width=(sos.extrapolation_width(float(t), M),) * 3) for t in ts]
        eng = sos.SosEngine(mu, tau, coefs, [0, L], sos._lib.SURFACE_NONE)
        eng.set_phase([P])
        assert eng.gensrc_enabled == (flag == "1")
        Cc = np.zeros((S, 2, N))
        Cc[:, 0] = 0.95 * P0
        I1 = eng.first_order(Cc)
        res = eng.solve(I1)
        torch.cuda.synchronize()
        out[flag] = (eng.to_host(res.I), res.n_orders.copy())
        eng.close()
    assert np.array_equal(out["1"][1], out["0"][1])
    assert relmax(out["1"][0], out["0"][0]) < 1e-12
    # ... and one of them against the oracle's single-layer functions
    k = 5
    P0, P = sos.phase_matrices("rayleigh", M, mu, 0.6, 0.0)
    A = so.contraction_matrix(P, mu, 0.95)
    In = so.I1_NumInt(tau[k], mu, float(ts[k]), 0.6, P0, 0.95, M)
    I = In.copy()
    n = 1
    ones = np.ones_like(I)
    while so.convergence_ratio(In if n > 1 else ones, I, M) >= 1e-4:
        n += 1
        In = so.In_NumInt(n, In @ A, In, tau[k], mu, float(ts[k]), 0.6, P, 0.95, M, method="recurrence")
        I = I + In
    assert n == int(out["1"][1][k])
    assert relmax(out["1"][0][k], I) < TOL


def test_wide_blend_falls_back_to_stored_sources(sos, monkeypatch):
    """A source large enough to push the find-first blend past the 128 upward columns whose raw I_n is kept next to
    mu = 0+ (the threshold of SOS_Aer_I1_In.py:103 is absolute): the plan must notice, switch to stored sources and
    return their result."""
    import torch
    L, M, S = 48, 801, 8
    N = 2 * M
    mu = sos.mu_grid(M)
    tau = np.tile(np.linspace(0, 0.3, L), (S, 1))
    P0, P = sos.phase_matrices("rayleigh", M, mu, 0.5, 0.0)
    w = sos.extrapolation_width(0.3, M)
    out = {}
    for flag in ("1", "0"):
        monkeypatch.setenv("SOS_B200_GENSRC", flag)
        coefs = [sos.ScenarioCoefficients(mu0=0.5, grd_alb=0.0, tauStar_tot=0.3, coef_atm=1.0, extrap_width=(w, w, w)) for _ in range(S)]
        eng = sos.SosEngine(mu, tau, coefs, [0, L], sos._lib.SURFACE_NONE)
        eng.set_phase([P])
        Cc = np.zeros((S, 2, N))
        Cc[:, 0] = 4.0e4 * P0 * (1.0 + 0.1 * np.arange(S))[:, None]     # a very bright sun: the absolute threshold bites
        I1 = eng.first_order(Cc)
        res = eng.solve(I1, max_orders=4)
        torch.cuda.synchronize()
        out[flag] = (eng.to_host(res.I), res.n_orders.copy(), eng.gensrc_enabled)
        eng.close()
    assert out["1"][2] is False, "the wide blend should have switched the plan to stored sources"
    assert np.array_equal(out["1"][1], out["0"][1])
    assert relmax(out["1"][0], out["0"][0]) < 1e-13


def test_critical_albedo_sweep_reports_the_forcing_it_evaluated(sos):
    """critical_albedo_sweep: the forcing returned for every point is the forcing of a stand-alone pair of solves at the
    omega it returned (whenever the sweep stopped on a midpoint it evaluated), and the bisection brackets the sign change."""
    pts = [sos.Scenario(nb_layers=80, nb_angles=41, mu0=m, tauStar_atm=0.124, tauStar_aer=t, grd_alb=r,
                        atm_phase=("rayleigh", 0.0), aer_phase=("hg", 0.6))
           for m in (0.4, 0.8) for t in (0.05, 0.3) for r in (0.05, 0.3)]
    omega, forcing, n_solves, evaluated = sos.critical_albedo_sweep(pts, width=0.05, return_evaluated=True)
    assert omega.shape == (8,) and np.all((omega > 0) & (omega < 1)) and n_solves >= 16

    def forcing_at(i, w):
        a = sos.solve_scenarios([dataclasses.replace(pts[i], alb_aer=float(w))])[0].toa_net_flux
        b = sos.solve_scenarios([dataclasses.replace(pts[i], tauStar_aer=0.0, alb_aer=1.0)])[0].toa_net_flux
        return a - b

    assert np.all(np.isfinite(forcing)) and np.all(np.isfinite(evaluated))
    for i in range(len(pts)):
        # `forcing[i]` is the forcing at `evaluated[i]`, the last omega the sweep solved for this point
        assert abs(forcing_at(i, evaluated[i]) - forcing[i]) < 1e-10 * max(1.0, abs(forcing[i])), i
        assert abs(omega[i] - evaluated[i]) <= 0.05
    for i in (0, 5):
        f_lo, f_hi = forcing_at(i, 0.01), forcing_at(i, 0.99)
        if f_lo * f_hi < 0:
            a, b = forcing_at(i, max(omega[i] - 0.06, 0.0)), forcing_at(i, min(omega[i] + 0.06, 1.0))
            assert a * b <= 0 or min(abs(a), abs(b)) < 2e-3
