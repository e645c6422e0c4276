"""GPU parity: the CUDA path, called through the reference-facing API / C ABI, against
(a) golden fixtures produced by the unmodified reference and (b) the CPU oracle on the same inputs.

FP64 tolerance stated by BASELINE.json north_star: 1e-10 relative on per-order radiances, fluxes
and heating rates, same scattering order at convergence.  Metric (hard part 4):
max|delta| / max|ref| per field, plus elementwise relative error where |ref| > 1e-6 max|ref|.
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
    sos_b200._lib.load()  # fails loudly if the extension is missing
    return sos_b200


@pytest.fixture(scope="module")
def so():
    import sos_oracle
    return sos_oracle


def smooth_source(tau, mu, tauStar):
    t, m = tau[:, None], mu[None, :]
    return (1 + 0.5 * m + 0.3 * m * m) * np.exp(-t / 0.5) * (1 + 0.2 * np.sin(3 * t / tauStar)) + 0.01


# ----------------------------------------------------------------------------------------------
# single-layer drop-in functions
# ----------------------------------------------------------------------------------------------
def test_single_layer_vs_golden(sos, golden):
    d = golden("single_layer.npz")
    for ci in range(int(d["ncases"])):
        L, M, ts, mu0, alb = d[f"c{ci}_params"]
        L, M = int(L), int(M)
        rows = d[f"c{ci}_rows"]
        mu = sos.mu_grid(M)
        tau = np.linspace(0, ts, L)
        P0, P = sos.phase_matrices(str(d[f"c{ci}_phase"]), M, mu, mu0, 0.5)
        I1 = sos.I1_NumInt(tau, mu, ts, mu0, P0, alb, M)
        assert I1.shape == (L, 2 * M) and I1.dtype == np.float64 and I1.flags["C_CONTIGUOUS"]
        assert relmax(I1[rows], d[f"c{ci}_I1"]) < TOL, ci
        assert relelem(I1[rows], d[f"c{ci}_I1"]) < TOL, ci
        J2 = sos.Jn_NumInt(2, I1, tau, mu, ts, mu0, P, alb, M)
        assert relmax(J2[rows], d[f"c{ci}_J2"]) < TOL, ci
        I2 = sos.In_NumInt(2, J2, I1, tau, mu, ts, mu0, P, alb, M, 0, 0)
        assert relmax(I2[rows], d[f"c{ci}_I2"]) < TOL, ci
        Is = sos.In_NumInt(2, smooth_source(tau, mu, ts), I1, tau, mu, ts, mu0, P, alb, M, 0, 0)
        assert relmax(Is[rows], d[f"c{ci}_Is"]) < TOL, ci
        assert tuple(d[f"c{ci}_mu12"]) == tuple(sos.mu_approx_In(mu, M))


@pytest.mark.parametrize("L,M,ts", [(37, 64, 0.4), (130, 101, 3.0), (257, 333, 9.0), (64, 512, 30.0 * 64 / 10000)])
def test_single_layer_vs_oracle_ragged(sos, so, L, M, ts):
    """Ragged sizes (N not a multiple of any tile) and several tau* regimes against the oracle."""
    mu = sos.mu_grid(M)
    tau = np.linspace(0, ts, L) ** 1.0
    tau[L // 2:] += 0.3 * ts / L  # a kink, as in SURVEY.md 8(d)
    mu0, alb = 0.6, 0.93
    P0, P = sos.phase_matrices("hg", M, mu, mu0, 0.7)
    a = sos.I1_NumInt(tau, mu, ts, mu0, P0, alb, M)
    b = so.I1_NumInt(tau, mu, ts, mu0, P0, alb, M)
    assert relmax(a, b) < TOL
    Ja = sos.Jn_NumInt(2, b, tau, mu, ts, mu0, P, alb, M)
    Jb = b @ so.contraction_matrix(P, mu, alb)
    assert relmax(Ja, Jb) < TOL
    Js = smooth_source(tau, mu, ts)
    Ia = sos.In_NumInt(2, Js, b, tau, mu, ts, mu0, P, alb, M, 0, 0)
    Ib = so.In_NumInt(2, Js, b, tau, mu, ts, mu0, P, alb, M, method="recurrence")
    assert relmax(Ia, Ib) < TOL


def test_blend_overrun_raises_index_error(sos):
    """Non-smooth input: the reference raises IndexError (SOS_Aer_I1_In.py:103); so do we."""
    L, M, ts = 40, 64, 0.5
    mu = sos.mu_grid(M)
    tau = np.linspace(0, ts, L)
    rng = np.random.default_rng(0)
    J = rng.standard_normal((L, 2 * M)) * 10
    with pytest.raises(IndexError):
        sos.In_NumInt(2, J, J, tau, mu, ts, 0.5, None, 1.0, M, 0, 0)


def test_inputs_not_mutated(sos):
    L, M, ts = 50, 101, 0.5
    mu = sos.mu_grid(M)
    tau = np.linspace(0, ts, L)
    P0, P = sos.phase_matrices("rayleigh", M, mu, 0.5)
    I1 = sos.I1_NumInt(tau, mu, ts, 0.5, P0, 1.0, M)
    keep = (I1.copy(), P.copy(), tau.copy(), mu.copy())
    J = sos.Jn_NumInt(2, I1, tau, mu, ts, 0.5, P, 1.0, M)
    Jk = J.copy()
    sos.In_NumInt(2, J, I1, tau, mu, ts, 0.5, P, 1.0, M, 0, 0)
    assert np.array_equal(I1, keep[0]) and np.array_equal(P, keep[1])
    assert np.array_equal(tau, keep[2]) and np.array_equal(mu, keep[3]) and np.array_equal(J, Jk)


# ----------------------------------------------------------------------------------------------
# three-region drivers
# ----------------------------------------------------------------------------------------------
def _run_driver(sos, kind, kw, keep):
    fn = sos.SOS_Aer_main_lambertian if kind == "lambertian" else sos.SOS_Aer_main_specular
    return fn(keep_orders=keep, atm_phase=("rayleigh", 0.5), aer_phase=("hg", 0.5), **kw)


@pytest.mark.parametrize("kind", ["specular", "lambertian"])
@pytest.mark.parametrize("tag", ["thin", "thick", "mu0hit"])
def test_small_drivers_vs_golden(sos, golden, kind, tag):
    d = golden("drivers_small.npz")
    key = f"{kind}_{tag}"
    kw = ast.literal_eval(str(d[key + "_kw"]))
    n_ref = int(d[key + "_n"])
    r = _run_driver(sos, kind, kw, keep=n_ref)
    assert r.n == n_ref
    assert np.array_equal(r.tau, d[key + "_tau"])
    assert [r.idx_up, r.idx_down] == list(d[key + "_idx"])
    assert relmax(r.I, d[key + "_I"]) < TOL
    for j, oid in enumerate(d[key + "_order_ids"]):
        assert relmax(r.I_saved[oid], d[key + "_orders"][j]) < TOL, oid
    assert relmax(r.flux_up, d[key + "_flux_up"]) < TOL
    assert relmax(r.flux_down, d[key + "_flux_down"]) < TOL
    assert relmax(r.net_flux, d[key + "_net_flux"]) < TOL
    assert relmax(r.diffusivity, d[key + "_diffusivity"]) < TOL
    # heating rate = difference of nearly equal fluxes / dz: compare against the flux scale
    scale = np.max(np.abs(d[key + "_flux_down"])) / abs(r.z_profile[1] - r.z_profile[0]) / (1.225 * 1004)
    assert np.max(np.abs(r.heating_rate - d[key + "_heating_rate"])) < TOL * scale


def test_more_than_sixteen_phase_functions_in_one_call(sos):
    """A phase sweep over 20 HG asymmetry factors on one grid (a plan holds 16 operands): solved in sub-batches, every
    member equal to its stand-alone solve."""
    scs = [sos.Scenario(nb_layers=96, nb_angles=101, tauStar_atm=0.124, tauStar_aer=0.12, grd_alb=0.15, alb_aer=0.97,
                        aer_phase=("hg", 0.2 + 0.03 * i)) for i in range(20)]
    out = sos.solve_scenarios(scs)
    for i in (0, 7, 16, 19):
        one = sos.solve_scenarios([scs[i]])[0]
        assert out[i].n == one.n and relmax(out[i].I, one.I) < 1e-12, i


def test_lambert_readme_mode_invariants_and_oracle(sos, so):
    """surface="lambert_readme": the n >= 2 Lambert coupling as README.md:215 states it (a named physics mode, SURVEY 8c;
    the shipped code has the opposite sign and drops [-h, 0], Q5).  Validated by invariants -- the surface returns the
    fraction rho of every order's downward irradiance, isotropically and with a positive radiance -- and against the
    oracle's restatement of the same formula."""
    kw = dict(nb_layers=120, nb_angles=101, mu0=0.6, tauStar_atm=0.3, tauStar_aer=0.4, alb_aer=0.95, grd_alb=0.4)
    sc = sos.Scenario(atm_phase=("rayleigh", 0.0), aer_phase=("hg", 0.6), surface="lambert_readme", **kw)
    r = sos.solve_scenarios([sc], keep_orders=6)[0]
    M, mu, L = sc.nb_angles, r.mu, sc.nb_layers
    for n in range(1, min(6, r.n)):                      # orders 2 ..
        In = r.I_saved[n]
        down = -so.trapz(In[L - 1, :M] * mu[:M], mu[:M])              # downward irradiance of this order at the surface (> 0)
        seed = In[L - 1, M + 40:]                                      # away from the mu -> 0+ blend: the raw surface radiance
        assert down > 0 and np.all(seed > 0)
        assert np.max(np.abs(seed - seed[0])) < 1e-13 * abs(seed[0])   # isotropic
        assert abs(seed[0] - 2 * sc.grd_alb * down) < 1e-12 * abs(seed[0])   # R = 2 rho F_down  <=>  F_up = rho F_down
        up = so.trapz(In[L - 1, M:] * mu[M:], mu[M:])
        assert abs(up - sc.grd_alb * down) < 2e-2 * up                 # (the blend next to mu = 0+ reshapes a few columns)
    # as coded, the same scenario has the opposite sign (Q5)
    rc = sos.solve_scenarios([dataclasses.replace(sc, surface="lambert")], keep_orders=2)[0]
    assert np.all(rc.I_saved[1][L - 1, M + 40:] < 0)
    # the oracle's restatement of the README formula
    P0a, Pa = sos.phase_matrices("rayleigh", M, mu, sc.mu0, 0.0)
    P0e, Pe = sos.phase_matrices("hg", M, mu, sc.mu0, 0.6)
    ref = so.solve(so.Scenario(surface="lambert_readme", **kw), P0a, Pa, P0e, Pe, method="recurrence", use_gemm=True)
    assert r.n == ref["n"] and relmax(r.I, ref["I"]) < TOL


def test_graphe_shim_reproduces_the_reference_curves(sos, so, golden):
    """The reference's output layer with its own signatures (sos.graphe_*, SOS_Aer_graphe.py:6,37,68,118,152): the curves
    the unmodified reference computed inside its graphe_* functions (captured by make_golden.py), reproduced from the
    golden radiance field through the device quadratures; per-order diffusivity against the NumPy formula."""
    d = golden("drivers_small.npz")
    key = "specular_thick"
    kw = ast.literal_eval(str(d[key + "_kw"]))
    r = _run_driver(sos, "specular", kw, keep=3)
    I, mu, z, tau = d[key + "_I"], r.mu, r.z_profile, d[key + "_tau"]
    L, M = I.shape[0], I.shape[1] // 2
    mu0, alb = kw.get("mu0", 0.5), kw.get("grd_alb", 0.0)
    F0 = np.pi / mu0
    assert relmax(sos.graphe_diffusivity(I, mu, z, L, "hg"), d[key + "_diffusivity"]) < TOL
    assert relmax(sos.graphe_flux(I, mu, z, L, M, tau, mu0, F0, alb, "hg"), d[key + "_net_flux"]) < TOL
    up, down = sos.graphe_flux_up_down(I, mu, z, L, M, tau, mu0, F0, alb, "hg")
    assert relmax(up, d[key + "_flux_up"]) < TOL and relmax(down, d[key + "_flux_down"]) < TOL
    hr = sos.graphe_heating_rate(I, mu, z, L, M, r.idx_up, r.idx_down, F0, mu0, tau, alb, "hg")
    scale = np.max(np.abs(d[key + "_flux_down"])) / abs(z[1] - z[0]) / (1.225 * 1004)
    assert np.max(np.abs(hr - d[key + "_heating_rate"])) < TOL * scale
    # another F0 only rescales the direct terms (the reference passes F0 explicitly)
    up2, down2 = sos.graphe_flux_up_down(I, mu, z, L, M, tau, mu0, 2.0 * F0, alb, "hg")
    assert relmax(down2 - down, -F0 * np.exp(-tau / mu0)) < 1e-9
    # per-order diffusivity, all saved orders in one launch
    dif = sos.graphe_successive_dif(r.I_saved, mu, z, L, M, "hg")
    assert dif.shape == (len(r.I_saved), L)
    for j, In in enumerate(r.I_saved):
        assert relmax(dif[j], so.diffusivity(In, mu)) < TOL, j
    with pytest.raises(ValueError):
        sos.graphe_successive_dif([I[:-1]], mu, z, L, M, "hg")


@pytest.mark.parametrize("tag", ["eva_spec", "eva_lamb", "thin_spec", "mixed_spec", "tau2_lamb"])
def test_n1002_drivers_vs_golden(sos, golden, tag):
    d = golden("drivers_n1002.npz")
    kw = ast.literal_eval(str(d[tag + "_kw"]))
    kind = kw.pop("kind")
    n_ref = int(d[tag + "_n"])
    r = _run_driver(sos, kind, kw, keep=n_ref)
    assert r.n == n_ref
    rows = d[tag + "_rows"]
    assert relmax(r.I[rows], d[tag + "_I_rows"]) < TOL
    assert relmax(r.I[::10], d[tag + "_I_sub"]) < TOL
    for j in range(n_ref):
        assert relmax(r.I_saved[j][rows], d[tag + "_order_rows"][j]) < TOL, j
        assert abs(np.sum(r.I_saved[j]) - d[tag + "_order_sum"][j]) < TOL * np.sum(np.abs(r.I_saved[j]))
    assert relmax(r.flux_up, d[tag + "_flux_up"]) < TOL and relmax(r.flux_down, d[tag + "_flux_down"]) < TOL
    assert relmax(r.net_flux, d[tag + "_net_flux"]) < TOL and relmax(r.diffusivity, d[tag + "_diffusivity"]) < TOL


DEFAULT_RUNS = {
    "eva_spec": ("specular", dict(tauStar_atm=0.124, grd_alb=0.15, alb_aer=0.97)),
    "eva_lamb": ("lambertian", dict(tauStar_atm=0.124, grd_alb=0.15, alb_aer=0.97)),
    "wildfire_lamb": ("lambertian", dict(tauStar_atm=0.124, tauStar_aer=0.0075, z_up=15, z_down=14, grd_alb=0.15, alb_aer=0.97)),
    "shipped_spec": ("specular", {}),
}


@pytest.mark.parametrize("tag,fold", [(t, "1") for t in DEFAULT_RUNS] + [("eva_spec", "0"), ("wildfire_lamb", "0")])
def test_default_grid_vs_golden(sos, golden, tag, fold, monkeypatch):
    """BASELINE configs 1-3 at the reference's 800 x 1002 grid (HG g=0.5 standing in for Mie), through the folded
    contraction (the default) and, for two of them, through the general kernel (SOS_B200_FOLD=0)."""
    monkeypatch.setenv("SOS_B200_FOLD", fold)
    d = golden(f"default_{tag}.npz")
    kind, kw = DEFAULT_RUNS[tag]
    n_ref = int(d["n"])
    r = _run_driver(sos, kind, kw, keep=n_ref)
    assert r.n == n_ref
    assert [r.idx_up, r.idx_down] == list(d["idx"])
    assert np.array_equal(r.tau, d["tau"])
    rows = d["rows"]
    assert relmax(r.I[rows], d["I_rows"]) < TOL and relelem(r.I[rows], d["I_rows"]) < 1e-9
    assert relmax(r.I[::40], d["I_sub"]) < TOL
    for j in range(n_ref):
        assert relmax(r.I_saved[j][rows], d["order_rows"][j]) < TOL, j
        assert abs(np.max(np.abs(r.I_saved[j])) - d["order_max"][j]) < TOL * d["order_max"][j]
    assert relmax(r.flux_up, d["flux_up"]) < TOL and relmax(r.flux_down, d["flux_down"]) < TOL
    assert relmax(r.net_flux, d["net_flux"]) < TOL and relmax(r.diffusivity, d["diffusivity"]) < TOL
    scale = np.max(np.abs(d["flux_down"])) / abs(r.z_profile[1] - r.z_profile[0]) / (1.225 * 1004)
    assert np.max(np.abs(r.heating_rate - d["heating_rate"])) < TOL * scale
    if tag == "eva_spec":
        assert abs(r.I[0, 751] - 0.24262858024159764) < 1e-12
        assert abs(r.flux_up[0] - 0.45982557863744034) < 1e-11


def test_thick_fwc_to_convergence(sos, golden):
    """Config-4 stand-in on a reduced grid: 100+ orders, same order count as the reference."""
    d = golden("thick_fwc.npz")
    L, M, ts, mu0, alb = d["params"]
    L, M = int(L), int(M)
    mu = sos.mu_grid(M)
    tau = np.linspace(0, ts, L)
    P0, P = sos.phase_matrices("fwc", M, mu, mu0)
    w = sos.extrapolation_width(ts, M)
    coef = sos.ScenarioCoefficients(mu0=mu0, grd_alb=0.0, tauStar_tot=ts, coef_atm=alb, extrap_width=(w, w, w))
    eng = sos.SosEngine(mu, tau[None], [coef], [0, L], sos._lib.SURFACE_NONE)
    eng.set_phase([P])
    Cc = np.zeros((1, 2, 2 * M))
    Cc[0, 0] = alb * P0
    I1 = eng.first_order(Cc)
    res = eng.solve(I1, keep_orders=60)
    assert int(res.n_orders[0]) == int(d["n"])
    I = eng.to_host(res.I)
    assert relmax(I[::25], d["I_sub"]) < TOL
    assert relmax(I[0], d["I_toa"]) < TOL and relmax(I[L - 1], d["I_surf"]) < TOL
    for n in (2, 3, 10, 50):
        got = res.orders[n - 2][:, : 2 * M].cpu().numpy()
        assert relmax(got[::50], d[f"order{n}_sub"]) < TOL, n
    eng.close()


# ----------------------------------------------------------------------------------------------
# size-independent properties at full size
# ----------------------------------------------------------------------------------------------
def test_batch_equals_single_and_chunking_is_invisible(sos):
    """Scenario batching and the scan chunk length are implementation details: results must not move."""
    base = dict(nb_layers=300, nb_angles=251, atm_phase=("rayleigh", 0.5), aer_phase=("hg", 0.6))
    scs = [sos.Scenario(mu0=0.5, tauStar_aer=0.12, alb_aer=0.97, grd_alb=0.15, **base),
           sos.Scenario(mu0=0.8, tauStar_aer=0.4, alb_aer=0.85, grd_alb=0.3, **base),
           sos.Scenario(mu0=0.3, tauStar_aer=0.02, alb_aer=1.0, grd_alb=0.05, **base)]
    batch = sos.solve_scenarios(scs)
    for i, sc in enumerate(scs):
        single = sos.solve_scenarios([sc])[0]
        assert single.n == batch[i].n
        assert relmax(batch[i].I, single.I) < 1e-13
    for chunk in (16, 37, 128):
        bs = sos.BatchSolver(scs, chunk_rows=chunk)
        res = bs.solve()
        out = bs.results(res)
        for i in range(len(scs)):
            assert out[i].n == batch[i].n
            assert relmax(out[i].I, batch[i].I) < 1e-12, chunk
        bs.engine.close()


def test_source_contraction_linearity_full_size(sos):
    """J is linear in I_{n-1} (config-4 width N = 1024, ragged rows): J(a x + b y) = a J(x) + b J(y)."""
    import torch
    L, M, ts = 1500, 512, 4.5
    mu = sos.mu_grid(M)
    tau = np.linspace(0, ts, L)
    P0, P = sos.phase_matrices("hg", M, mu, 0.5, 0.8)
    rng = np.random.default_rng(1)
    x = rng.random((L, 2 * M))
    y = rng.random((L, 2 * M))
    Jx = sos.Jn_NumInt(2, x, tau, mu, ts, 0.5, P, 0.9, M)
    Jy = sos.Jn_NumInt(2, y, tau, mu, ts, 0.5, P, 0.9, M)
    Jz = sos.Jn_NumInt(2, 2.0 * x - 0.5 * y, tau, mu, ts, 0.5, P, 0.9, M)
    assert relmax(Jz, 2.0 * Jx - 0.5 * Jy) < 1e-13
    # and against a float64 matmul on the same device (torch is the checker here, not the product)
    d = np.diff(mu)
    w = np.zeros_like(mu)
    w[:-1] += d / 2
    w[1:] += d / 2
    A = (0.9 / 4) * (w[:, None] * P[:, ::-1].T)
    ref = (torch.as_tensor(x).cuda() @ torch.as_tensor(np.ascontiguousarray(A)).cuda()).cpu().numpy()
    assert relmax(Jx, ref) < 1e-13


def test_mu_block_sharding_matches_unsharded_single_gpu(sos):
    """mu-block sharding (config 4): four plans on one GPU, each owning a 256-column block, must
    reproduce the unsharded order bit for bit (same arithmetic per column) -- the NCCL all-gather
    between them is covered by the gloo test and tools/mu_shard_check.py."""
    import torch
    L, M, ts, mu0, alb = 400, 512, 1.2, 0.5, 0.9
    N = 2 * M
    mu = sos.mu_grid(M)
    tau = np.linspace(0, ts, L)
    P0, P = sos.phase_matrices("hg", M, mu, mu0, 0.8)
    w = sos.extrapolation_width(ts, M)
    coef = [sos.ScenarioCoefficients(mu0=mu0, grd_alb=0.0, tauStar_tot=ts, coef_atm=alb, extrap_width=(w, w, w))]
    Cc = np.zeros((1, 2, N))
    Cc[0, 0] = alb * P0

    def one_order(cols=None, fold=False):
        eng = sos.SosEngine(mu, tau[None], coef, [0, L], sos._lib.SURFACE_NONE, fold=fold)
        eng.set_phase([P])
        assert eng.folded == fold
        if cols is not None:
            eng.set_columns(*cols)
        I1 = eng.first_order(Cc)
        eng.reset(I1)
        I = I1.clone()
        J = eng.source(I1)
        In = eng.sweeps(J, accumulate_into=I)
        torch.cuda.synchronize()
        out = (J[:, :N].cpu().numpy(), In[:, :N].cpu().numpy(), I[:, :N].cpu().numpy(), eng.ratios().cpu().numpy())
        eng.close()
        return out

    Jf, Inf, If, rf = one_order()
    zone_lo = M - w - 5
    blocks = sos.mu_blocks(N, M, 4, zone_lo)
    rmax = np.full_like(rf, -np.inf)
    for c0, c1 in blocks:
        Jb, Inb, Ib, rb = one_order((c0, c1))
        assert np.array_equal(Jb[:, c0:c1], Jf[:, c0:c1])
        assert np.array_equal(Inb[:, c0:c1], Inf[:, c0:c1])
        assert np.array_equal(Ib[:, c0:c1], If[:, c0:c1])
        rmax = np.maximum(rmax, rb)
    assert np.array_equal(rmax, rf)
    # the folded contraction (default for full-column plans) is the same sum reassociated
    Jd, Ind, Id, rd = one_order(fold=True)
    assert relmax(Jd, Jf) < 1e-13 and relmax(Ind, Inf) < 1e-12 and relmax(Id, If) < 1e-13
    # a boundary inside the mu -> 0 zone is refused
    eng = sos.SosEngine(mu, tau[None], coef, [0, L], sos._lib.SURFACE_NONE)
    with pytest.raises(sos.SosError):
        eng.set_columns(0, 500)
    eng.close()


def test_radiative_forcing_and_critical_albedo(sos, golden):
    """SOS_Aer_radiative_forcing / SOS_Aer_critical_albedo drop-ins against the reference's own numbers,
    then the corrected batched sweep (genuine tau_aer = 0 baseline)."""
    d = golden("forcing.npz")
    for tag in ("a", "b"):
        c = ast.literal_eval(str(d[tag + "_case"]))
        L, M, mu0 = c["L"], c["M"], c["mu0"]
        mu = sos.mu_grid(M)
        P0a, Pa = sos.phase_matrices("rayleigh", M, mu, mu0)
        P0h, Ph = sos.phase_matrices("hg", M, mu, mu0, 0.5)
        tau = sos.tau_profile(c["ta"], c["te"], 120, c["z_up"], c["z_down"], L)
        _, iu, idn = sos.aerosol_rows(120, c["z_up"], c["z_down"], L)
        dtau_aer, dtau_atm, F0 = c["te"] / (idn + 1 - iu), c["ta"] / L, np.pi / mu0
        args = (dtau_aer, c["ta"], dtau_atm, Ph, P0h, c["alb_aer"], Pa, P0a, 1.0, c["rho"], F0, mu, mu0, M, tau, L, iu, idn)
        ref = float(d[tag + "_toa_net_flux"])
        got = sos.SOS_Aer_radiative_forcing(0, *args, tauStar_tot=c["ta"] + c["te"])
        assert abs(got - ref) < TOL * abs(ref)
        assert sos.SOS_Aer_radiative_forcing(c["te"], *args, tauStar_tot=c["ta"] + c["te"]) == float(d[tag + "_forcing"]) == 0.0
        crit = sos.SOS_Aer_critical_albedo(c["te"], dtau_aer, c["ta"], dtau_atm, Ph, P0h, Pa, P0a, 1.0, c["rho"], F0, mu, mu0,
                                           M, tau, L, iu, idn, tauStar_tot=c["ta"] + c["te"])
        assert crit == float(d[tag + "_critical"]) == 0.5
    # (the corrected batched sweep is checked in test_gpu_workload.py::test_critical_albedo_sweep_reports_the_forcing_it_evaluated)


@pytest.mark.parametrize("name,g", [("rayleigh", 0.0), ("hg", 0.5), ("hg", 0.75), ("fwc", 0.0)])
def test_device_phase_builder_vs_reference(sos, golden, name, g):
    """sos_build_phase against the reference's own builders (fixtures) and the host builder."""
    d = golden("phase_small.npz")
    for M in (21, 41):
        for mu0 in (0.5, 0.8):
            mu = sos.mu_grid(M)
            coef = [sos.ScenarioCoefficients(mu0=mu0, grd_alb=0.0, tauStar_tot=1.0, coef_atm=1.0)]
            eng = sos.SosEngine(mu, np.linspace(0, 1, 8)[None], coef, [0, 8], sos._lib.SURFACE_NONE)
            P, P0 = eng.build_phase_matrix(name, g, mu0=mu0)
            P, P0 = P.cpu().numpy(), P0.cpu().numpy()
            eng.close()
            key = f"{name}_M{M}_mu0{mu0}_g{g}"
            assert relmax(P, d[key + "_P"]) < 1e-12 and relmax(P0, d[key + "_P0"]) < 1e-12
    # full size against the host builder (itself pinned on the reference at M = 501 by the n1002 fixture)
    M = 501
    mu = sos.mu_grid(M)
    w = sos.extrapolation_width(1.0, M)
    coef = [sos.ScenarioCoefficients(mu0=0.5, grd_alb=0.0, tauStar_tot=1.0, coef_atm=1.0, extrap_width=(w, w, w))]
    eng = sos.SosEngine(mu, np.linspace(0, 1, 8)[None], coef, [0, 8], sos._lib.SURFACE_NONE)
    P, P0 = eng.build_phase_matrix(name, g, mu0=0.5)
    Ph = sos.phase_P(name, M, mu, g) if name != "fwc" else None
    if Ph is not None:
        assert relmax(P.cpu().numpy(), Ph) < 1e-12
    assert relmax(P0.cpu().numpy(), sos.phase_P0(name, M, mu, 0.5, g)) < 1e-12
    eng.close()


def _oracle_solve(so, sos, sc, atm=("rayleigh", 0.0), aer=("hg", 0.5)):
    M = sc.nb_angles
    mu = so.mu_grid(M)
    P0a, Pa = sos.phase_matrices(atm[0], M, mu, sc.mu0, atm[1])
    P0e, Pe = sos.phase_matrices(aer[0], M, mu, sc.mu0, aer[1])
    osc = so.Scenario(mu0=sc.mu0, z0=sc.z0, z_up=sc.z_up, z_down=sc.z_down, nb_layers=sc.nb_layers,
                      tauStar_atm=sc.tauStar_atm, tauStar_aer=sc.tauStar_aer, grd_alb=sc.grd_alb, alb_atm=sc.alb_atm,
                      alb_aer=sc.alb_aer, nb_angles=M, surface=sc.surface, threshold=sc.threshold)
    return so.solve(osc, P0a, Pa, P0e, Pe, method="recurrence", use_gemm=True, keep_orders=False)


def test_edge_scenarios_vs_oracle(sos, so):
    """Edge cases the fixtures do not reach, one mixed batch per surface kind, against the oracle:
    no aerosol at all (second contraction coefficient exactly 0), absorbing gas (omega_atm < 1), black and
    mirror surfaces, mu0 off the grid; then an optically thick column (tau_ref > 4: extrapolation width
    class 0.06 M in the lower regions, 0.02 M above; 33 orders)."""
    base = dict(nb_layers=96, nb_angles=251, atm_phase=("rayleigh", 0.0), aer_phase=("hg", 0.5))
    for surface in ("specular", "lambert"):
        scs = [sos.Scenario(mu0=0.5, tauStar_atm=0.124, tauStar_aer=0.0, grd_alb=0.15, surface=surface, **base),
               sos.Scenario(mu0=0.37, tauStar_atm=0.6, tauStar_aer=0.9, alb_aer=0.8, alb_atm=0.9, grd_alb=0.3, surface=surface, **base),
               sos.Scenario(mu0=0.9, tauStar_atm=0.3, tauStar_aer=0.2, alb_aer=1.0, grd_alb=0.0, surface=surface, **base),
               sos.Scenario(mu0=0.2, tauStar_atm=0.05, tauStar_aer=0.01, alb_aer=0.95, grd_alb=1.0, surface=surface,
                            z_up=60.0, z_down=30.0, **base)]
        # scenarios 0-2 share the aerosol rows, scenario 3 has its own geometry -> two batches inside
        got = sos.solve_scenarios(scs)
        for sc, r in zip(scs, got):
            ref = _oracle_solve(so, sos, sc)
            assert r.n == ref["n"], (surface, sc.mu0, r.n, ref["n"])
            assert relmax(r.I, ref["I"]) < TOL, (surface, sc.mu0)
    thick = sos.Scenario(nb_layers=400, nb_angles=151, mu0=0.37, tauStar_atm=0.8, tauStar_aer=3.6, alb_aer=0.8, alb_atm=0.9,
                         grd_alb=0.3, z_up=100.0, z_down=10.0, surface="specular", atm_phase=("rayleigh", 0.0),
                         aer_phase=("hg", 0.5))
    r = sos.solve_scenarios([thick])[0]
    ref = _oracle_solve(so, sos, thick)
    assert r.n == ref["n"] and relmax(r.I, ref["I"]) < TOL


def test_taylor_columns_three_regions_vs_oracle(sos, so):
    """M = 1201: |mu| = 1/1200 < 1e-3 is a Taylor column, ten windowed columns (SOS_Aer_In_limit.py:79-107)."""
    sc = sos.Scenario(nb_layers=64, nb_angles=1201, mu0=0.5, tauStar_atm=0.03, tauStar_aer=0.02, grd_alb=0.4, alb_aer=0.9,
                      atm_phase=("iso", 0.0), aer_phase=("iso", 0.0), surface="specular")
    r = sos.solve_scenarios([sc])[0]
    ref = _oracle_solve(so, sos, sc, atm=("iso", 0.0), aer=("iso", 0.0))
    assert r.n == ref["n"]
    assert relmax(r.I, ref["I"]) < TOL


def test_long_blend_beyond_the_row_zone(sos, so):
    """The find-first threshold is absolute (SOS_Aer_I1_In.py:103): a large source pushes the blend index
    far from mu = 0+, past the 128 columns the row-wise kernel keeps in shared memory."""
    L, M, ts = 48, 801, 0.3
    mu = sos.mu_grid(M)
    tau = np.linspace(0, ts, L)
    Js = 100.0 * smooth_source(tau, mu, ts)
    ref_idx = np.zeros(L, dtype=int)
    lay = so.Layout(tau=tau, mu=mu, nb_angles=M, regions=[(0, L)], tau_ref=[ts], thick=False, surface="none", grd_alb=0.0)
    ref = so.order_sweeps(lay, Js, method="recurrence", blend_index_out=ref_idx)
    assert ref_idx.max() > M + 140, ref_idx.max()      # the case really leaves the zone
    got = sos.In_NumInt(2, Js, Js, tau, mu, ts, 0.5, None, 1.0, M, 0, 0)
    assert relmax(got, ref) < TOL


def test_large_batch_tile_shape_matches_single_solves(sos):
    """A batch large enough to switch the contraction to its 128 x 144 tile shape (and to pack aerosol
    rows of many scenarios into shared two-operand tiles) must reproduce single-scenario solves, which use
    64 x 128 tiles with split operand passes."""
    base = dict(nb_layers=800, nb_angles=501, atm_phase=("rayleigh", 0.0))
    scs = []
    for k in range(30):
        scs.append(sos.Scenario(mu0=0.2 + 0.025 * k, tauStar_atm=0.124, tauStar_aer=0.01 + 0.015 * k,
                                alb_aer=0.75 + 0.008 * k, grd_alb=(0.05, 0.15, 0.3)[k % 3],
                                aer_phase=(("hg", 0.5), ("hg", 0.7))[k % 2], surface="specular", **base))
    batch = sos.solve_scenarios(scs, quadratures=False, fold=False)
    for i in (0, 13, 29):
        single = sos.solve_scenarios([scs[i]], quadratures=False, fold=False)[0]
        assert single.n == batch[i].n
        assert relmax(batch[i].I, single.I) < 1e-13, i
    # ... and the folded contraction (centrosymmetric operands: half the multiply-adds) the general one
    folded = sos.solve_scenarios(scs, quadratures=False)
    for i in range(len(scs)):
        assert folded[i].n == batch[i].n
        assert relmax(folded[i].I, batch[i].I) < 1e-12, i


def test_folded_contraction_vs_general_and_asymmetric_operand(sos):
    """The folded kernel (u = x[k] + x[N-1-k], v = x[k] - x[N-1-k], two M x M operands) against the general one on
    random rows: odd and even M, ragged sizes, two-operand aerosol rows in packed and split tiles.  An operand that
    is not centrosymmetric must keep the general kernel."""
    import torch
    rng = np.random.default_rng(7)
    for L, M, S in ((70, 37, 1), (131, 100, 3), (64, 501, 2), (300, 129, 40)):
        N = 2 * M
        mu = sos.mu_grid(M)
        iu, idn = L // 3, L // 3 + 9
        tau = np.tile(np.linspace(0, 0.7, L), (S, 1))
        _, Pa = sos.phase_matrices("rayleigh", M, mu, 0.5, 0.0)
        _, Pe = sos.phase_matrices("hg", M, mu, 0.5, 0.8)
        w = sos.extrapolation_width(0.7, M)
        coefs = [sos.ScenarioCoefficients(mu0=0.5, grd_alb=0.1, tauStar_tot=0.7, coef_atm=0.9 + 0.01 * s, coef_mix_atm=0.3 + 0.02 * s,
                                          coef_mix_aer=0.6 - 0.01 * s, phase_atm=0, phase_aer=1, extrap_width=(w, w, w))
                 for s in range(S)]
        x = rng.random((S, L, N)) * np.exp(2.0 * rng.standard_normal((S, L, N)))
        out = {}
        for fold in (False, True):
            eng = sos.SosEngine(mu, tau, coefs, [0, iu, idn + 1, L], sos._lib.SURFACE_SPECULAR, fold=fold)
            eng.set_phase([Pa, Pe])
            assert eng.folded == fold
            if fold:
                assert eng.fold_defect < 1e-13
            J = eng.source(eng.to_field(x))
            torch.cuda.synchronize()
            out[fold] = eng.to_host(J).reshape(S, L, N)
            eng.close()
        assert relelem(out[True], out[False]) < 1e-11, (L, M, S)
        assert relmax(out[True], out[False]) < 1e-13, (L, M, S)
    # asymmetric operand: stays on the general kernel
    L, M = 40, 64
    mu = sos.mu_grid(M)
    _, P = sos.phase_matrices("hg", M, mu, 0.5, 0.6)
    P = P * (1.0 + 1e-6 * rng.random(P.shape))
    w = sos.extrapolation_width(0.5, M)
    eng = sos.SosEngine(mu, np.linspace(0, 0.5, L)[None], [sos.ScenarioCoefficients(mu0=0.5, grd_alb=0.0, tauStar_tot=0.5,
                        coef_atm=1.0, extrap_width=(w, w, w))], [0, L], sos._lib.SURFACE_NONE)
    eng.set_phase([P])
    assert not eng.folded and eng.fold_defect > 1e-8
    eng.close()


def test_folded_kernel_variants_agree(sos, monkeypatch):
    """The folded contraction's internal variants are the same sums: premixed aerosol operand (one pass) vs two operand
    passes with a rescaled accumulator, transformer warps vs per-warp u/v -- on packed tiles (S = 20) and on the split
    tiles of a single scenario."""
    import torch
    rng = np.random.default_rng(11)
    L, M = 120, 165
    N = 2 * M
    mu = sos.mu_grid(M)
    iu, idn = 40, 66                      # 27 aerosol rows: 4 segments, the last one ragged
    _, Pa = sos.phase_matrices("rayleigh", M, mu, 0.5, 0.0)
    _, Pe = sos.phase_matrices("hg", M, mu, 0.5, 0.7)
    w = sos.extrapolation_width(0.7, M)
    for S in (20, 1):
        tau = np.tile(np.linspace(0, 0.7, L), (S, 1))
        coefs = [sos.ScenarioCoefficients(mu0=0.5, grd_alb=0.1, tauStar_tot=0.7, coef_atm=0.9 + 0.004 * s, coef_mix_atm=0.3 + 0.01 * s,
                                          coef_mix_aer=(0.0 if s == 3 else 0.6 - 0.01 * s), phase_atm=0, phase_aer=1, extrap_width=(w, w, w))
                 for s in range(S)]
        x = rng.random((S, L, N)) * np.exp(rng.standard_normal((S, L, N)))
        out = {}
        for premix, xform in (("1", "1"), ("0", "1"), ("1", "0"), ("0", "0")):
            monkeypatch.setenv("SOS_FOLD_PREMIX", premix)
            monkeypatch.setenv("SOS_FOLD_XFORM", xform)
            eng = sos.SosEngine(mu, tau, coefs, [0, iu, idn + 1, L], sos._lib.SURFACE_SPECULAR, fold=True)
            eng.set_phase([Pa, Pe])
            assert eng.folded
            J = eng.source(eng.to_field(x))
            torch.cuda.synchronize()
            out[(premix, xform)] = eng.to_host(J).reshape(S, L, N)
            eng.close()
        ref = out[("1", "1")]
        for key, val in out.items():
            assert relmax(val, ref) < 1e-13, (S, key)
            assert relelem(val, ref) < 1e-11, (S, key)


def test_lognormal_mie_aerosol_device_build_and_solve(sos, so):
    """The reference's EVA / wildfire aerosols (log-normal Mie mixtures, README.md:90-111) through the host Mie stand-in:
    the 6001-point mixture table goes to the device phase builder (same tabulated family as the FWC cloud), and a
    three-region solve with that aerosol reproduces the oracle fed with the same P0 / P."""
    M = 151
    mu = sos.mu_grid(M)
    w = sos.extrapolation_width(1.0, M)
    coef = [sos.ScenarioCoefficients(mu0=0.5, grd_alb=0.0, tauStar_tot=1.0, coef_atm=1.0, extrap_width=(w, w, w))]
    for aerosol in (sos.EVA_AEROSOL, sos.WILDFIRE_AEROSOL):
        eng = sos.SosEngine(mu, np.linspace(0, 1, 8)[None], coef, [0, 8], sos._lib.SURFACE_NONE)
        P, P0 = eng.build_phase_matrix("mie_lognormal", aerosol, mu0=0.5)
        P0h, Ph = sos.phase_matrices("mie_lognormal", M, mu, 0.5, aerosol)
        assert relmax(P.cpu().numpy(), Ph) < 1e-12 and relmax(P0.cpu().numpy(), P0h) < 1e-12
        eng.close()
    for name, aerosol, kw in (("eva", sos.EVA_AEROSOL, sos.EVA), ("wildfire", sos.WILDFIRE_AEROSOL, sos.WILDFIRE)):
        sc = sos.Scenario(nb_layers=160, nb_angles=M, mu0=0.5, surface="specular", atm_phase=("rayleigh", 0.0),
                          aer_phase=("mie_lognormal", aerosol), **kw)
        r = sos.solve_scenarios([sc])[0]
        ref = _oracle_solve(so, sos, sc, aer=("mie_lognormal", aerosol))
        assert r.n == ref["n"], name
        assert relmax(r.I, ref["I"]) < TOL, name


def test_lowrank_operand_rows_match_the_dense_contraction(sos, monkeypatch):
    """Rayleigh (rank 2) and isotropic (rank 1) operands: the rows that use them alone are contracted as (I Us) Vt by
    jn_lowrank_kernel; HG stays dense.  Same J as the dense folded kernel on random rows, ragged segments, S = 1 and a batch."""
    import torch
    rng = np.random.default_rng(5)
    for L, M, S, atm in ((131, 100, 3, "rayleigh"), (70, 37, 1, "iso"), (200, 251, 24, "rayleigh")):
        N = 2 * M
        mu = sos.mu_grid(M)
        iu, idn = L // 3, L // 3 + 10
        tau = np.tile(np.linspace(0, 0.7, L), (S, 1))
        _, Pa = sos.phase_matrices(atm, M, mu, 0.5, 0.0)
        _, Pe = sos.phase_matrices("hg", M, mu, 0.5, 0.8)
        w = sos.extrapolation_width(0.7, M)
        coefs = [sos.ScenarioCoefficients(mu0=0.5, grd_alb=0.1, tauStar_tot=0.7, coef_atm=0.9 + 0.003 * s, coef_mix_atm=0.3 + 0.01 * s,
                                          coef_mix_aer=0.6 - 0.01 * s, phase_atm=0, phase_aer=1, extrap_width=(w, w, w))
                 for s in range(S)]
        x = rng.random((S, L, N)) * np.exp(rng.standard_normal((S, L, N)))
        out = {}
        for flag in ("1", "0"):
            monkeypatch.setenv("SOS_B200_LOWRANK", flag)
            eng = sos.SosEngine(mu, tau, coefs, [0, iu, idn + 1, L], sos._lib.SURFACE_SPECULAR, fold=True)
            eng.set_phase([Pa, Pe])
            assert eng.folded
            assert eng.lowrank == ([2 if atm == "rayleigh" else 1, 0] if flag == "1" else [0, 0])
            J = eng.source(eng.to_field(x))
            torch.cuda.synchronize()
            out[flag] = eng.to_host(J).reshape(S, L, N)
            eng.close()
        assert relmax(out["1"], out["0"]) < 1e-13, (L, M, S)
        assert relelem(out["1"], out["0"]) < 1e-11, (L, M, S)
    # single homogeneous layer (the Jn_NumInt case): every row is low rank, no dense tile at all
    L, M = 90, 64
    mu = sos.mu_grid(M)
    _, P = sos.phase_matrices("rayleigh", M, mu, 0.5, 0.0)
    x = rng.random((L, 2 * M))
    tau = np.linspace(0, 0.5, L)
    monkeypatch.setenv("SOS_B200_LOWRANK", "1")
    sos.clear_cache()
    J1 = sos.Jn_NumInt(2, x, tau, mu, 0.5, 0.5, P, 0.9, M)
    monkeypatch.setenv("SOS_B200_FOLD", "0")
    sos.clear_cache()
    J0 = sos.Jn_NumInt(2, x, tau, mu, 0.5, 0.5, P, 0.9, M)
    sos.clear_cache()
    assert relmax(J1, J0) < 1e-13


@pytest.mark.parametrize("surface", ["specular", "lambert"])
def test_first_order_entry_points_agree_bit_for_bit(sos, surface):
    """sos_first_order_tab (coefficient planes assembled on the device from the table of distinct solar phase vectors) and
    sos_first_order2 (planes assembled on the host, SOS_Aer_main_specular.py:52-53) produce the same bits, and the second
    destination of the first order (the accumulator of the order loop) holds the same field."""
    import torch
    scen = [sos.Scenario(nb_layers=90, nb_angles=75, tauStar_atm=0.124, tauStar_aer=0.03 + 0.05 * i, mu0=(0.2, 0.5, 0.9)[i % 3],
                         alb_aer=0.8 + 0.02 * i, grd_alb=0.1 * (i % 4), atm_phase=("rayleigh", 0.0),
                         aer_phase=(("hg", 0.5), ("hg", 0.7), ("fwc", 0.0))[i % 3], surface=surface) for i in range(9)]
    bs = sos.BatchSolver(scen)
    eng = bs.engine
    assert eng.table_fits(bs.P0tab.shape[0], len(scen), eng.N) and bs.P0tab.shape[0] < 2 * len(scen)
    acc = eng.new_field(zero=True)
    a = bs.first_order(also_into=acc).clone()
    b = eng.first_order(bs.Ccoef)
    torch.cuda.synchronize()
    assert torch.equal(a, b)
    assert torch.equal(acc[:, : eng.N], a[:, : eng.N])
    # a single scenario with two phase functions does not fit the table's staging: the planes go up as they are
    one = sos.BatchSolver(scen[:1])
    assert not one.engine.table_fits(one.P0tab.shape[0], 1, one.engine.N)
    r1 = one.results(one.solve(), quadratures=False)[0]
    rb = bs.results(bs.solve(), quadratures=False)[0]
    assert r1.n == rb.n and relmax(r1.I, rb.I) < 1e-12
    one.engine.close()
    bs.engine.close()
