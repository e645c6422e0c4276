"""CPU-only checks of the host side: grid helpers, phase builders, the C-ABI surface."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import sos_oracle as so
from conftest import ROOT, relmax

import sos_b200 as sos
from importlib import import_module

G = import_module("sos-radiative-transfer_b200.grid")


def test_library_exports_every_declared_symbol():
    """include/sos_b200.h <-> ctypes prototypes <-> symbols of libsos_b200.so."""
    hdr = open(os.path.join(ROOT, "include", "sos_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(sos_[a-z0-9_]+)\s*\(", hdr))
    assert declared, "no declarations parsed"
    assert declared == set(sos._lib.SIGNATURES), declared ^ set(sos._lib.SIGNATURES)
    lib = sos._lib.load()
    for name in declared:
        assert hasattr(lib, name), name
    assert lib.sos_abi_version() == 1


def test_abi_host_only_entry_points():
    lib = sos._lib.load()
    assert lib.sos_strerror(0) == b"ok"
    assert b"invalid" in lib.sos_strerror(-1)
    idx, ns, off = (C.c_int * 4)(), (C.c_int * 4)(), (C.c_int * 4)()
    total = lib.sos_extrap_layout(501, idx, ns, off)
    assert list(idx) == [2, 10, 20, 30] and list(ns) == [2, 5, 5, 5]
    assert total == 2 * 2 + 10 * 5 + 20 * 5 + 30 * 5
    assert total == G.extrapolation_tables(sos.mu_grid(501), 501).size
    total = lib.sos_extrap_layout(201, idx, ns, off)
    assert list(idx) == [1, 4, 8, 12] and list(ns) == [2, 4, 5, 5]
    assert total == G.extrapolation_tables(sos.mu_grid(201), 201).size
    # invalid arguments are reported as error codes, never as exceptions or crashes
    assert lib.sos_plan_create(None, None, None, None, None, None, 0) == -1
    assert lib.sos_source(None, None, None, None) == -1
    assert lib.sos_plan_destroy(None) == 0


def test_no_cpu_fallback_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(sos.SosError):
        sos.I1_NumInt(np.linspace(0, 1, 10), sos.mu_grid(8), 1.0, 0.5, np.ones(16), 1.0, 8)


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "sos-radiative-transfer_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert "sos_oracle" not in text and "ref_harness" not in text, f
                assert not re.search(r"^\s*(from|import)\s+oracle", text, flags=re.M), f


@pytest.mark.parametrize("kw", [dict(tauStar_atm=0.104, tauStar_aer=0.12, z0=120, z_up=25, z_down=17, nb_layers=800),
                                dict(tauStar_atm=0.124, tauStar_aer=0.0075, z0=120, z_up=15, z_down=14, nb_layers=800),
                                dict(tauStar_atm=0.5, tauStar_aer=1.5, z0=120, z_up=17, z_down=25, nb_layers=70)])
def test_tau_profile_and_rows(kw, golden):
    a = sos.tau_profile(**kw)
    zu, zd = max(kw["z_up"], kw["z_down"]), min(kw["z_up"], kw["z_down"])
    b = so.tau_profile(kw["tauStar_atm"], kw["tauStar_aer"], kw["z0"], zu, zd, kw["nb_layers"])
    assert np.array_equal(a, b)
    if kw["nb_layers"] == 800 and kw["z_up"] == 25:
        d = golden("default_shipped_spec.npz")
        assert np.array_equal(a, d["tau"])
        _, iu, idn = sos.aerosol_rows(kw["z0"], kw["z_up"], kw["z_down"], 800)
        assert [iu, idn] == list(d["idx"]) == [633, 686]
        assert a[-1] == kw["tauStar_atm"] + kw["tauStar_aer"]


def test_mu_grid_and_mu_approx():
    for M in (64, 101, 501, 1201):
        mu = sos.mu_grid(M)
        assert np.array_equal(mu, so.mu_grid(M))
        assert sos.mu_approx_In(mu, M) == so.mu_approx_In(mu, M)
    with pytest.raises(IndexError):
        sos.mu_approx_In(np.concatenate((np.linspace(-1, 0, 8), np.linspace(0, 0.005, 8))), 8)


def test_extrapolation_tables_match_polyfit_semantics():
    for M in (101, 201, 251, 501, 512, 1201):
        mu = sos.mu_grid(M)
        flat = G.extrapolation_tables(mu, M)
        pos = 0
        for f in (0.005, 0.02, 0.04, 0.06):
            w = int(f * M)
            W, src0, ns = so.extrapolation_matrix(mu[:M], w)
            got = flat[pos: pos + W.size].reshape(W.shape)
            pos += W.size
            assert np.allclose(got, W, rtol=1e-12, atol=1e-12)
        assert pos == flat.size
    for tau_ref, M in ((0.05, 501), (0.0625, 501), (0.5, 501), (1.0, 501), (3.9, 501), (4.0, 501), (30, 512)):
        assert sos.extrapolation_width(tau_ref, M) == so.extrapolation_width(tau_ref, M)


def test_phase_builders_vs_reference(golden):
    d = golden("phase_small.npz")
    n = 0
    for k in d.files:
        if not k.endswith("_P"):
            continue
        name, M, mu0, g = k.split("_")[:4]
        M, mu0, g = int(M[1:]), float(mu0[3:]), float(g[1:])
        P0, P = sos.phase_matrices(name, M, sos.mu_grid(M), mu0, g)
        assert relmax(P, d[k]) < 1e-13 and relmax(P0, d[k[:-2] + "_P0"]) < 1e-13
        # the reference normalises every column to trapz = 4 (SOS_Aer_phase_func.py:131)
        assert np.allclose(so.trapz(P, sos.mu_grid(M), axis=0), 4.0, rtol=1e-13)
        n += 1
    assert n >= 16
    # full-size subsample pinned by the n1002 fixture
    d2 = golden("drivers_n1002.npz")
    M = 501
    P0a, Pa = sos.phase_matrices("rayleigh", M, sos.mu_grid(M), 0.5)
    P0h, Ph = sos.phase_matrices("hg", M, sos.mu_grid(M), 0.5, 0.5)
    assert relmax(Pa[::25, ::25], d2["phase_sub_atm"]) < 1e-13 and relmax(Ph[::25, ::25], d2["phase_sub_aer"]) < 1e-13
    assert relmax(P0a, d2["P0_atm"]) < 1e-13 and relmax(P0h, d2["P0_aer"]) < 1e-13


def test_scenario_defaults_are_the_shipped_literals():
    sc = sos.Scenario()
    # SOS_Aer_main_specular.py:23-57
    assert (sc.mu0, sc.z0, sc.z_up, sc.z_down, sc.nb_layers) == (0.5, 120, 25, 17, 800)
    assert (sc.tauStar_atm, sc.tauStar_aer, sc.grd_alb, sc.alb_atm, sc.alb_aer, sc.nb_angles) == (0.104, 0.120, 1, 1.0, 1.0, 501)


@pytest.mark.parametrize("name,g,M", [("rayleigh", 0.0, 37), ("hg", 0.8, 64), ("fwc", 0.0, 101)])
def test_folded_contraction_algebra(name, g, M):
    """The identity behind csrc/gemm_fold.cuh, restated in NumPy on the oracle's contraction operand: the operand
    of every reference phase function is centrosymmetric to rounding, and J[:, j], J[:, N-1-j] = u B+ +- v B- with
    u, v = x[k] +- x[N-1-k] reproduces the full contraction."""
    mu = so.mu_grid(M)
    N = 2 * M
    _, P = sos.phase_matrices(name, M, mu, 0.6, g)
    A = so.contraction_matrix(P, mu, 0.9)                     # A[k, m], SOS_Aer_I1_In.py:73
    defect = np.abs(A - A[::-1, ::-1]).max() / np.abs(A).max()
    assert defect < 1e-12                                     # engine.FOLD_DEFECT_MAX
    rng = np.random.default_rng(M)
    x = rng.random((9, N)) * np.exp(2.0 * rng.standard_normal((9, N)))
    J = x @ A
    As = 0.5 * (A + A[::-1, ::-1])                            # sos_build_folded symmetrises first
    mirror = As[:M, ::-1][:, :M]                              # A[k, N-1-j]
    Bp, Bm = 0.5 * (As[:M, :M] + mirror), 0.5 * (As[:M, :M] - mirror)
    xr = x[:, ::-1][:, :M]                                    # x[N-1-k]
    u, v = x[:, :M] + xr, x[:, :M] - xr
    Jf = np.empty_like(J)
    Jf[:, :M] = u @ Bp + v @ Bm
    Jf[:, ::-1][:, :M] = u @ Bp - v @ Bm
    assert np.max(np.abs(Jf - J) / np.abs(J)) < 1e-12
    # layout the C ABI promises for the folded operand
    rows, ld = C.c_int(), C.c_int()
    n = sos._lib.load().sos_fold_layout(M, C.byref(rows), C.byref(ld))
    assert rows.value % 16 == 0 and rows.value >= M and ld.value == 2 * rows.value and n == rows.value * ld.value


def test_mie_stand_in_against_published_values():
    """The host Lorenz-Mie series (mie.py, the stand-in for the reference's unpinned miepython) against published
    numbers: the test sphere of Bohren & Huffman's BHMIE listing (m = 1.55, x = 2 pi 0.525 / 0.6328), Wiscombe's
    MIEV0 cases m = 1.5 and m = 1.5 - 0.1i at x = 10, the Rayleigh limit, and the 'albedo' normalisation of
    i_unpolarized that the reference's mixing relies on (SOS_Aer_phase_func.py:419,693)."""
    M = sos.mie
    qext, qsca, qback, _ = M.efficiencies(1.55 + 0j, 2 * np.pi * 0.525 / 0.6328)
    assert abs(qext - 3.10543) < 5e-6 and abs(qsca - 3.10543) < 5e-6 and abs(qback - 2.92534) < 5e-6
    qext, qsca, _, g = M.efficiencies(1.5 + 0j, 10.0)
    assert abs(qext - 2.881999) < 2e-6 and abs(qsca - 2.881999) < 2e-6 and abs(g - 0.742913) < 2e-6
    qext, qsca, _, g = M.efficiencies(1.5 - 0.1j, 10.0)          # either sign of Im m means absorption
    assert abs(qext - 2.459791) < 2e-6 and abs(qsca - 1.235144) < 2e-6 and abs(g - 0.922350) < 2e-6
    assert M.efficiencies(1.5 + 0.1j, 10.0)[0] == qext
    x, m = 0.01, 1.5 + 0j
    assert abs(M.efficiencies(m, x)[1] / ((8 / 3) * x ** 4 * abs((m * m - 1) / (m * m + 2)) ** 2) - 1) < 1e-4
    mu_s = np.linspace(-1, 1, 20001)
    i_ray = M.i_unpolarized(m, x, mu_s)
    assert np.max(np.abs(i_ray / i_ray[10000] - (1 + mu_s ** 2))) < 2e-4
    for m, x in ((1.5 + 0j, 10.0), (1.7 + 0.03j, 2.0), (1.44 + 0j, 0.3)):
        qe, qs, _, g = M.efficiencies(m, x)
        i = M.i_unpolarized(m, x, mu_s)
        integral = 2 * np.pi * so.trapz(i, mu_s)
        assert abs(integral - qs / qe) < 5e-6                   # integral over 4 pi = single-scattering albedo
        assert abs(2 * np.pi * so.trapz(i * mu_s, mu_s) / integral - g) < 5e-6


def test_lognormal_mie_family_goes_through_the_tabulated_builder():
    """'mie_lognormal' = the reference's log_normal_mie mixture (as coded) as a 6001-point table, then the same
    azimuth average / normalisation as every other family (trapz(P[:, n], mu) = 4, trapz(P0, mu) = 2)."""
    xs, ys = sos.phase_table("mie_lognormal", sos.EVA_AEROSOL)
    assert xs.shape == (6001,) and xs[0] == -1 and xs[-1] == 1 and np.all(ys > 0)
    assert ys[-1] > 50 * ys[0]                                    # r_m = 0.5 um at 0.55 um: strongly forward peaked
    xw, yw = sos.phase_table("mie_lognormal", sos.WILDFIRE_AEROSOL)
    assert 3 < yw[-1] / yw[0] < 30                                # r_m = 0.065 um: much closer to Rayleigh
    Mg = 41
    mu = so.mu_grid(Mg)
    P0, P = sos.phase_matrices("mie_lognormal", Mg, mu, 0.5, sos.EVA_AEROSOL)
    assert np.allclose([so.trapz(P[:, n], mu) for n in range(2 * Mg)], 4.0, rtol=1e-13)
    assert abs(so.trapz(P0, mu) - 2.0) < 1e-13
    assert np.max(np.abs(P - P[::-1, ::-1])) < 1e-12 * np.max(P)  # centrosymmetric like every other family
    # the physical mixture (weights n(r) r^2 Qsca(x)) is available but is not what the reference computes
    _, yp = sos.mie.lognormal_table(*sos.EVA_AEROSOL, as_coded=False)
    assert np.max(np.abs(yp / yp.max() - ys / ys.max())) > 1e-3


def test_lowrank_structure_of_the_molecular_operand():
    """The property behind csrc/gemm_lowrank.cuh (sos_build_lowrank_mu2), restated in NumPy: every row of the Rayleigh
    operand of the reference is affine in mu_m^2 (isotropic: constant), so the factors read off two of its columns --
    alpha = A[:, M-1] (mu = 0), beta = A[:, 0] - A[:, M-1] (mu^2 = 1) -- reproduce it to rounding and (I Us) Vt = I A;
    HG, FWC and the Mie mixture do not have that structure and must be refused by the residual test."""
    M = 101
    N = 2 * M
    mu = so.mu_grid(M)
    rng = np.random.default_rng(3)
    x = rng.random((7, N)) * np.exp(rng.standard_normal((7, N)))
    resid = {}
    for name, g in (("rayleigh", 0.0), ("iso", 0.0), ("hg", 0.5), ("hg", 0.05), ("fwc", 0.0)):
        _, P = sos.phase_matrices(name, M, mu, 0.5, g)
        A = so.contraction_matrix(P, mu, 1.0)
        alpha, beta = A[:, M - 1], A[:, 0] - A[:, M - 1]
        fit = alpha[:, None] + beta[:, None] * (mu * mu)[None, :]
        resid[(name, g)] = np.max(np.abs(A - fit)) / np.max(np.abs(A))
        if name in ("rayleigh", "iso"):
            J = (x @ np.stack([alpha, beta], 1)) @ np.stack([np.ones(N), mu * mu], 0)
            assert np.max(np.abs(J - x @ A) / np.abs(x @ A)) < 1e-13, name
            assert (np.max(np.abs(beta)) <= 1e-15 * np.max(np.abs(A))) == (name == "iso")
            assert np.linalg.matrix_rank(A, tol=1e-13 * np.linalg.norm(A, 2)) == (1 if name == "iso" else 2)
    assert resid[("rayleigh", 0.0)] < 1e-14 and resid[("iso", 0.0)] < 1e-14          # engine.LOWRANK_RESIDUAL_MAX
    # even a weakly anisotropic HG is refused (the advisor's worry about truncating near-low-rank operands)
    assert resid[("hg", 0.5)] > 1e-3 and resid[("hg", 0.05)] > 1e-6 and resid[("fwc", 0.0)] > 1e-3


def test_reference_aerosol_names_resolve_to_the_mie_mixture():
    """'eva' / 'wildfire' (the names of the reference's phase_func dispatcher) are the log-normal Mie mixtures of mie.py."""
    from importlib import import_module
    D = import_module("sos-radiative-transfer_b200.drivers")
    M = 21
    mu = so.mu_grid(M)
    cache = D.PhaseCache()
    P0a, Pa, ka = cache.get(("eva", 0.0), M, mu, 0.5)
    P0b, Pb, kb = cache.get(("mie_lognormal", sos.EVA_AEROSOL), M, mu, 0.5)
    assert ka == kb == ("mie_lognormal", tuple(float(v) for v in sos.EVA_AEROSOL), M)
    assert np.array_equal(P0a, P0b) and np.array_equal(Pa, Pb)
    P0w, Pw = sos.phase_matrices("wildfire", M, mu, 0.5)
    P0x, Px = sos.phase_matrices("mie_lognormal", M, mu, 0.5, sos.WILDFIRE_AEROSOL)
    assert np.array_equal(P0w, P0x) and np.array_equal(Pw, Px)
    assert cache.get(("hg", 0.5), M, mu, 0.5)[2] == ("hg", 0.5, M)


def test_mie_table_disk_cache_is_keyed_on_every_parameter(tmp_path, monkeypatch):
    """The log-normal Mie table cache (the reference keys its .npy cache on nb_angles and a few scenario
    numbers only, SOS_Aer_global_va.py:17-83): one file per parameter set, identical values on reload, a changed
    refractive index never picks up a stale file, SOS_B200_CACHE=off writes nothing."""
    import sos_b200 as sos
    monkeypatch.setenv("SOS_B200_CACHE", str(tmp_path))
    sos.mie.lognormal_table.cache_clear()
    p = (0.55, 1.44, 0.0, 0.2, 1.3)
    mu_a, a = sos.mie.lognormal_table(*p)
    files = sorted(os.listdir(tmp_path))
    assert len(files) == 1 and files[0].startswith("mie_lognormal_") and files[0].endswith(".npy")
    sos.mie.lognormal_table.cache_clear()
    _, b = sos.mie.lognormal_table(*p)            # from disk
    assert np.array_equal(a, b) and len(os.listdir(tmp_path)) == 1
    sos.mie.lognormal_table.cache_clear()
    _, c = sos.mie.lognormal_table(0.55, 1.50, 0.0, 0.2, 1.3)
    assert len(os.listdir(tmp_path)) == 2 and not np.array_equal(a, c)
    # a corrupt file is rebuilt, not trusted
    path = os.path.join(tmp_path, files[0])
    with open(path, "wb") as f:
        f.write(b"not a npy file")
    sos.mie.lognormal_table.cache_clear()
    _, d = sos.mie.lognormal_table(*p)
    assert np.array_equal(a, d)
    sos.clear_caches(disk=True)
    assert os.listdir(tmp_path) == []
    monkeypatch.setenv("SOS_B200_CACHE", "off")
    sos.mie.lognormal_table(*p)
    assert os.listdir(tmp_path) == []
    sos.mie.lognormal_table.cache_clear()


def test_bench_arms_share_one_config():
    """The driver compares the `config` objects of the two bench arms: both come from bench.workload_config."""
    import ast as _ast
    import bench
    cfg = bench.workload_config(96)
    assert cfg["workload"] == bench.WORKLOAD and cfg["layers"] == 800 and cfg["mu_columns"] == 1002
    src = open(os.path.join(ROOT, "bench.py")).read()
    tree = _ast.parse(src)
    uses = [n for n in _ast.walk(tree) if isinstance(n, _ast.Call) and getattr(n.func, "id", "") == "workload_config"]
    assert len(uses) >= 2   # reference arm and our arm
    assert '"config": {"workload"' not in src.split("def thick_record")[0].split("def run_reference")[1].split("def ")[0]


def test_phase_sweeps_are_split_to_the_plan_limits():
    """More than 16 phase functions on one grid: solve_scenarios cuts the group into sub-batches a plan accepts."""
    from sos_b200 import drivers as D
    scs = [sos.Scenario(nb_layers=40, nb_angles=21, aer_phase=("hg", 0.3 + 0.02 * i)) for i in range(40)]
    parts = D.split_for_plan_limits(scs, list(range(40)))
    assert [i for p in parts for i in p] == list(range(40))
    for p in parts:
        specs = {scs[i].atm_phase for i in p} | {scs[i].aer_phase for i in p}
        assert len(specs) <= D.MAX_PHASES_PER_PLAN
    assert len(parts) == 3     # 1 atmosphere + 15 aerosols per plan
    assert D.split_for_plan_limits(scs[:5], [0, 1, 2, 3, 4]) == [[0, 1, 2, 3, 4]]


def test_batch_preparation_matches_the_per_scenario_formulas():
    """BatchSolver._prepare (host side of a batch: no device needed): batched tau profiles == grid.tau_profile row by row,
    and the table / rows / weights handed to sos_first_order_tab reproduce the first-order coefficient planes of
    SOS_Aer_main_specular.py:52-53,104-292 (C0 = alb_atm P0_atm, C1 = C0 f_atm + alb_aer P0_aer f_aer) bit for bit."""
    import bench
    drivers = import_module("sos_b200").drivers
    scen = bench.make_scenarios(sos, 24, rank=1, L=120, M=41)
    bs = object.__new__(sos.BatchSolver)
    bs._key = drivers._group_key(scen[0])
    L, M, bs.idx_up, bs.idx_down, _ = bs._key
    bs.L, bs.M, bs.N = L, M, 2 * M
    bs.mu = sos.grid.mu_grid(M)
    bs._phases = drivers.PhaseCache()
    bs._device_phase = False
    bs._mat_index, bs._mats, bs._mat_keys = {}, [], []
    for _ in range(2):                       # the second pass reuses the host buffers of the first
        coefs = bs._prepare(scen)
    C = bs.Ccoef
    assert bs.P0tab.shape[0] < 2 * len(scen) and sos.engine.SosEngine.table_fits(bs.P0tab.shape[0], len(scen), bs.N)
    for i, sc in enumerate(scen):
        tau = sos.grid.tau_profile(sc.tauStar_atm, sc.tauStar_aer, sc.z0, sc.z_up, sc.z_down, L)
        assert np.array_equal(bs.tau[i], tau)
        P0a, _ = sos.phase_matrices(sc.atm_phase[0], M, bs.mu, sc.mu0, sc.atm_phase[1])
        P0e, _ = sos.phase_matrices(sc.aer_phase[0], M, bs.mu, sc.mu0, sc.aer_phase[1])
        dtau_aer = sc.tauStar_aer / (bs.idx_down + 1 - bs.idx_up)
        dtau_atm = sc.tauStar_atm / L
        f_atm, f_aer = dtau_atm / (dtau_atm + dtau_aer), dtau_aer / (dtau_atm + dtau_aer)
        assert np.array_equal(C[i, 0], P0a * sc.alb_atm)
        assert np.array_equal(C[i, 1], (P0a * sc.alb_atm) * f_atm + (P0e * sc.alb_aer) * f_aer)
        assert coefs["tauStar_tot"][i] == sc.tauStar_atm + sc.tauStar_aer
        assert np.array_equal(bs.P0tab[bs.P0idx[i, 0]], P0a) and np.array_equal(bs.P0tab[bs.P0idx[i, 1]], P0e)
        # widths of the mu -> 0- extrapolation per region (SOS_Aer_main_specular.py:342-345) and the operands of the scenario
        assert coefs["extrap_width"][i, 0] == sos.extrapolation_width(float(tau[bs.idx_up - 1]), M)
        assert coefs["extrap_width"][i, 1] == coefs["extrap_width"][i, 2] == sos.extrapolation_width(float(tau[bs.idx_down]), M)
        for spec, col in ((sc.atm_phase, "phase_atm"), (sc.aer_phase, "phase_aer")):
            assert bs._mat_keys[coefs[col][i]][0] == spec[0]
        assert coefs["mu0"][i] == sc.mu0 and coefs["grd_alb"][i] == sc.grd_alb and coefs["coef_atm"][i] == sc.alb_atm
        assert coefs["coef_mix_atm"][i] == sc.alb_atm * f_atm and coefs["coef_mix_aer"][i] == sc.alb_aer * f_aer
    # every width class the thresholds distinguish, scalar vs batched
    for M_ in (21, 101, 501, 1201):
        ts = np.array([0.0, 0.01, 0.0625, 0.0626, 0.5, 1.0, 1.0001, 3.99, 4.0, 30.0])
        assert [int(w) for w in sos.grid.extrapolation_widths(ts, M_)] == [sos.extrapolation_width(float(t_), M_) for t_ in ts]


def test_bench_deals_the_job_evenly_over_the_ranks():
    """bench.make_scenarios(world > 1): the ranks' batches partition members 0 .. world*S - 1 of the sweep, every rank gets S
    of them and batches of the same make-up (cost-sorted serpentine rounds: the mean cost proxy differs by well under 1 %)."""
    import bench
    S, world = 96, 8
    job = bench.make_scenarios(sos, world * S, 0)
    key = lambda sc: (sc.tauStar_aer, sc.mu0, sc.alb_aer, sc.grd_alb, sc.aer_phase[0])
    dealt = [bench.make_scenarios(sos, S, r, world=world) for r in range(world)]
    assert all(len(b) == S for b in dealt)
    assert sorted(key(sc) for b in dealt for sc in b) == sorted(key(sc) for sc in job)
    means = [np.mean([bench.cost_proxy(sc) for sc in b]) for b in dealt]
    assert (max(means) - min(means)) / np.mean(means) < 0.01
    assert [key(sc) for sc in bench.make_scenarios(sos, S, 0, world=1)] == [key(sc) for sc in job[:S]]
    # the second deal of bench.py goes by measured orders to convergence: every rank gets the same distribution of them
    rng = np.random.default_rng(5)
    n_of = rng.integers(7, 24, size=world * S)
    job2, deals = bench.job_deal(sos, S, world, cost=n_of)
    assert sorted(i for d in deals for i in d) == list(range(world * S)) and [key(sc) for sc in job2] == [key(sc) for sc in job]
    hist = [np.bincount(n_of[d], minlength=24) for d in deals]
    assert max(np.abs(h - hist[0]).max() for h in hist) <= 2
    assert max(int(n_of[d].sum()) for d in deals) - min(int(n_of[d].sum()) for d in deals) <= 16
