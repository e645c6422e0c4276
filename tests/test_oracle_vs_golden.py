"""The oracle (oracle/sos_oracle.py) against fixtures produced by the UNMODIFIED reference.

CPU only.  Tolerance: the oracle restates the same arithmetic, so it must agree with the
reference far below the 1e-10 product tolerance; 1e-12 is asserted (observed <= 2e-14).
"""
import ast

import numpy as np
import pytest

import sos_oracle as so
from conftest import relmax

TOL = 1e-12


def _phase(name, M, mu, mu0, g=0.5):
    import sos_b200
    return sos_b200.phase_matrices(name, M, mu, mu0, g)


def test_single_layer_functions(golden):
    d = golden("single_layer.npz")
    for ci in range(int(d["ncases"])):
        L, M, ts, mu0, alb = d[f"c{ci}_params"]
        L, M = int(L), int(M)
        rows = d[f"c{ci}_rows"]
        mu = so.mu_grid(M)
        tau = np.linspace(0, ts, L)
        P0, P = _phase(str(d[f"c{ci}_phase"]), M, mu, mu0, 0.5)
        I1 = so.I1_NumInt(tau, mu, ts, mu0, P0, alb, M)
        assert relmax(I1[rows], d[f"c{ci}_I1"]) < TOL
        if M <= 251:
            J2 = so.Jn_NumInt(2, I1, tau, mu, ts, mu0, P, alb, M)
        else:
            J2 = I1 @ so.contraction_matrix(P, mu, alb)
        assert relmax(J2[rows], d[f"c{ci}_J2"]) < TOL
        for method in ("slices", "recurrence"):
            I2 = so.In_NumInt(2, J2, I1, tau, mu, ts, mu0, P, alb, M, method=method)
            assert relmax(I2[rows], d[f"c{ci}_I2"]) < TOL, (ci, method)
        assert tuple(d[f"c{ci}_mu12"]) == so.mu_approx_In(mu, M)


def test_contraction_matrix_is_the_trapezoid(golden):
    d = golden("single_layer.npz")
    L, M, ts, mu0, alb = d["c0_params"]
    L, M = int(L), int(M)
    mu = so.mu_grid(M)
    tau = np.linspace(0, ts, L)
    P0, P = _phase("hg", M, mu, mu0, 0.5)
    I1 = so.I1_NumInt(tau, mu, ts, mu0, P0, alb, M)
    a = so.Jn_NumInt(2, I1, tau, mu, ts, mu0, P, alb, M)
    b = I1 @ so.contraction_matrix(P, mu, alb)
    assert relmax(b, a) < 1e-13


@pytest.mark.parametrize("kind", ["specular", "lambertian"])
@pytest.mark.parametrize("tag", ["thin", "thick", "mu0hit"])
def test_small_drivers(golden, kind, tag):
    d = golden("drivers_small.npz")
    key = f"{kind}_{tag}"
    kw = ast.literal_eval(str(d[key + "_kw"]))
    M = kw["nb_angles"]
    sc = so.Scenario(surface="lambert" if kind == "lambertian" else "specular", **kw)
    mu = so.mu_grid(M)
    P0a, Pa = _phase("rayleigh", M, mu, sc.mu0)
    P0h, Ph = _phase("hg", M, mu, sc.mu0, 0.5)
    for method in ("slices", "recurrence"):
        res = so.solve(sc, P0a, Pa, P0h, Ph, method=method, use_gemm=(method == "recurrence"))
        assert res["n"] == int(d[key + "_n"])
        assert np.array_equal(res["tau"], d[key + "_tau"])
        assert [res["idx_up"], res["idx_down"]] == list(d[key + "_idx"])
        assert relmax(res["I"], d[key + "_I"]) < TOL
        for j, oid in enumerate(d[key + "_order_ids"]):
            assert relmax(res["I_saved"][oid], d[key + "_orders"][j]) < TOL, (method, oid)
    # quadratures
    I, tau, z = res["I"], res["tau"], res["z"]
    F0 = np.pi / sc.mu0
    up, down = so.flux_up_down(I, mu, M, tau, sc.mu0, F0, sc.grd_alb)
    assert relmax(up, d[key + "_flux_up"]) < TOL and relmax(down, d[key + "_flux_down"]) < TOL
    assert relmax(so.net_flux(I, mu, tau, sc.mu0, F0, sc.grd_alb), d[key + "_net_flux"]) < TOL
    assert relmax(so.diffusivity(I, mu), d[key + "_diffusivity"]) < 1e-11
    hr = so.heating_rate(I, mu, z, M, res["idx_up"], res["idx_down"], F0, sc.mu0, tau, sc.grd_alb)
    assert relmax(hr, d[key + "_heating_rate"]) < 1e-9  # differences of nearly equal fluxes


@pytest.mark.parametrize("tag", ["eva_spec", "eva_lamb", "thin_spec", "mixed_spec", "tau2_lamb"])
def test_n1002_drivers(golden, tag):
    """M = 501: windowed columns, extrapolation widths 2/10/20."""
    d = golden("drivers_n1002.npz")
    kw = ast.literal_eval(str(d[tag + "_kw"]))
    kind = kw.pop("kind")
    M = 501
    sc = so.Scenario(surface="lambert" if kind == "lambertian" else "specular", nb_angles=M, **kw)
    mu = so.mu_grid(M)
    P0a, Pa = _phase("rayleigh", M, mu, 0.5)
    P0h, Ph = _phase("hg", M, mu, 0.5, 0.5)
    res = so.solve(sc, P0a, Pa, P0h, Ph, method="recurrence", use_gemm=True)
    assert res["n"] == int(d[tag + "_n"])
    rows = d[tag + "_rows"]
    assert relmax(res["I"][rows], d[tag + "_I_rows"]) < TOL
    assert relmax(res["I"][::10], d[tag + "_I_sub"]) < TOL
    for j in range(res["n"]):
        assert relmax(res["I_saved"][j][rows], d[tag + "_order_rows"][j]) < TOL, j
    F0 = np.pi / sc.mu0
    up, down = so.flux_up_down(res["I"], mu, M, res["tau"], sc.mu0, F0, sc.grd_alb)
    assert relmax(up, d[tag + "_flux_up"]) < TOL and relmax(down, d[tag + "_flux_down"]) < TOL


@pytest.mark.parametrize("tag", ["mu01", "mu1", "mid"])
def test_fwc_aerosol_three_region_driver(golden, tag):
    """FWC cloud as the AEROSOL of the three-region specular driver (a third of bench.py's scenarios), at the corners
    of its sweep: mu0 = 0.1 / 1.0 (on the grid), omega_aer = 0.7 / 1.0 -- fixture from the unmodified reference."""
    d = golden("drivers_fwc3.npz")
    kw = ast.literal_eval(str(d[tag + "_kw"]))
    M = kw["nb_angles"]
    sc = so.Scenario(surface="specular", **kw)
    mu = so.mu_grid(M)
    P0a, Pa = _phase("rayleigh", M, mu, sc.mu0)
    P0f, Pf = _phase("fwc", M, mu, sc.mu0)
    assert relmax(P0f, d[tag + "_P0_aer"]) < TOL
    res = so.solve(sc, P0a, Pa, P0f, Pf, method="recurrence", use_gemm=True)
    assert res["n"] == int(d[tag + "_n"])
    rows = d[tag + "_rows"]
    assert relmax(res["I"][rows], d[tag + "_I_rows"]) < TOL
    assert relmax(res["I"][::5], d[tag + "_I_sub"]) < TOL
    for j in range(res["n"]):
        assert relmax(res["I_saved"][j][rows], d[tag + "_order_rows"][j]) < TOL, j
    F0 = np.pi / sc.mu0
    up, down = so.flux_up_down(res["I"], mu, M, res["tau"], sc.mu0, F0, sc.grd_alb)
    assert relmax(up, d[tag + "_flux_up"]) < TOL and relmax(down, d[tag + "_flux_down"]) < TOL


def test_default_grid_eva_specular(golden):
    """BASELINE config 3 at the reference's default 800 x 1002 grid (HG g=0.5 Mie stand-in)."""
    d = golden("default_eva_spec.npz")
    M = 501
    sc = so.Scenario(surface="specular", tauStar_atm=0.124, grd_alb=0.15, alb_aer=0.97)
    mu = so.mu_grid(M)
    P0a, Pa = _phase("rayleigh", M, mu, 0.5)
    P0h, Ph = _phase("hg", M, mu, 0.5, 0.5)
    res = so.solve(sc, P0a, Pa, P0h, Ph, method="recurrence", use_gemm=True)
    assert res["n"] == int(d["n"]) == 9
    assert [res["idx_up"], res["idx_down"]] == list(d["idx"]) == [633, 686]
    assert relmax(res["I"][d["rows"]], d["I_rows"]) < TOL
    assert relmax(res["I"][::40], d["I_sub"]) < TOL
    # SURVEY.md Appendix B.3 spot values
    assert abs(res["I"][0, 751] - 0.24262858024159764) < 1e-13
    assert abs(res["I"][799, 250] - 0.2855181943549694) < 1e-13


def test_thick_fwc_first_orders(golden):
    """Config-4 stand-in: thick single FWC layer; first 10 orders (the full 100+ are a GPU test)."""
    d = golden("thick_fwc.npz")
    L, M, ts, mu0, alb = d["params"]
    L, M = int(L), int(M)
    mu = so.mu_grid(M)
    tau = np.linspace(0, ts, L)
    P0, P = _phase("fwc", M, mu, mu0)
    assert relmax(P0, d["P0"]) < TOL and relmax(P[::8, ::8], d["P_sub"]) < TOL
    A = so.contraction_matrix(P, mu, alb)
    In = so.I1_NumInt(tau, mu, ts, mu0, P0, alb, M)
    I = In.copy()
    ratios = []
    for n in range(2, 11):
        ratios.append(so.convergence_ratio(In if n > 2 else np.ones_like(I), I, M))
        In = so.In_NumInt(n, In @ A, In, tau, mu, ts, mu0, P, alb, M, method="recurrence")
        I = I + In
        if f"order{n}_sub" in d.files:
            assert relmax(In[::50], d[f"order{n}_sub"]) < TOL, n
    assert np.allclose(ratios, d["ratios"][: len(ratios)], rtol=1e-11, atol=0)


def test_toa_net_flux_vs_reference_forcing(golden):
    """SOS_Aer_radiative_forcing (SOS_Aer_critical_albedo.py:20-389): TOA net flux of the solve; the
    shipped forcing is identically 0 and the shipped bisection returns 0.5 (Q19)."""
    d = golden("forcing.npz")
    for tag in ("a", "b"):
        c = ast.literal_eval(str(d[tag + "_case"]))
        M = c["M"]
        mu = so.mu_grid(M)
        P0a, Pa = _phase("rayleigh", M, mu, c["mu0"])
        P0h, Ph = _phase("hg", M, mu, c["mu0"], 0.5)
        sc = so.Scenario(mu0=c["mu0"], nb_layers=c["L"], nb_angles=M, tauStar_atm=c["ta"], tauStar_aer=c["te"],
                         grd_alb=c["rho"], alb_aer=c["alb_aer"], z_up=c["z_up"], z_down=c["z_down"])
        res = so.solve(sc, P0a, Pa, P0h, Ph, method="recurrence", use_gemm=True)
        got = so.toa_net_flux(res["I"], mu, M, res["tau"], c["mu0"], np.pi / c["mu0"], c["rho"])
        assert abs(got - float(d[tag + "_toa_net_flux"])) < 1e-12 * abs(float(d[tag + "_toa_net_flux"]))
        assert float(d[tag + "_forcing"]) == 0.0 and float(d[tag + "_critical"]) == 0.5
