"""Generate the golden fixtures in this directory by running the UNMODIFIED reference.

Only usable where /root/reference exists (the build container):

    python tests/golden/make_golden.py small      # seconds..minutes, small grids
    python tests/golden/make_golden.py n1002      # ~10 min: M=501 grids, reduced L
    python tests/golden/make_golden.py default    # ~25 min: the 800 x 1002 scenarios
    python tests/golden/make_golden.py thick      # ~10 min: thick single-layer FWC
    python tests/golden/make_golden.py forcing    # SOS_Aer_radiative_forcing / critical albedo, small grids
    python tests/golden/make_golden.py fwc_table  # the FWC data table (input data)
    python tests/golden/make_golden.py fwc3       # ~5 min: FWC cloud as the AEROSOL of the three-region driver

Every array written here is an output of reference code (imported from where it
lies through oracle/ref_harness.py); no reference source is copied.  Big fields
are subsampled so the fixtures stay small; the scenario parameters stored next
to them are enough to regenerate the inputs with the package's own builders.
Reference phase matrices at M=501 take ~100 s each, so they are cached under
/tmp/sos_golden (scratch, not committed).
"""
from __future__ import annotations

import os
import sys
import time
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
warnings.simplefilter("ignore")

import ref_harness as rh  # noqa: E402

CACHE = "/tmp/sos_golden"
os.makedirs(CACHE, exist_ok=True)


def mu_grid(M):
    return np.concatenate((np.linspace(-1, 0, M), np.linspace(0, 1, M)))


def ref_phase(name, M, mu0, g=0.5):
    path = os.path.join(CACHE, f"P_{name}_M{M}_mu0{mu0}_g{g}.npz")
    if os.path.exists(path):
        d = np.load(path)
        return d["P0"], d["P"]
    P0, P = rh.phase_matrices(name, M, mu_grid(M), mu0, g)
    np.savez(path, P0=P0, P=P)
    return P0, P


def smooth_source(tau, mu, tauStar):
    """an analytic source the reference digests without IndexError."""
    t = tau[:, None]
    m = mu[None, :]
    return (1 + 0.5 * m + 0.3 * m * m) * np.exp(-t / 0.5) * (1 + 0.2 * np.sin(3 * t / tauStar)) + 0.01


def capture_quadratures(I, out, M, mu0, grd_alb, aer="hg"):
    """Run the reference's graphe_* functions and capture their local arrays."""
    g = rh.load_reference()["graphe"]
    res = {}

    def grab(fn, names, *args):
        got = {}

        def prof(frame, event, arg):
            if event == "return" and frame.f_code.co_name == fn.__name__:
                got.update(frame.f_locals)
        import contextlib, io
        cwd = os.getcwd()
        os.chdir(rh.load_reference()["scratch"])
        try:
            with contextlib.redirect_stdout(io.StringIO()):
                sys.setprofile(prof)
                try:
                    fn(*args)
                finally:
                    sys.setprofile(None)
        finally:
            os.chdir(cwd)
        return {n: np.array(got[n]) for n in names}

    mu, z, tau = out["mu"], out["z_profile"], out["tau"]
    L = I.shape[0]
    F0 = np.pi / mu0
    res["diffusivity"] = grab(g.graphe_diffusivity, ["dif"], I, mu, z, L, aer)["dif"]
    res["net_flux"] = grab(g.graphe_flux, ["flux"], I, mu, z, L, M, tau, mu0, F0, grd_alb, aer)["flux"]
    q = grab(g.graphe_flux_up_down, ["flux_up", "flux_down"], I, mu, z, L, M, tau, mu0, F0, grd_alb, aer)
    res["flux_up"], res["flux_down"] = q["flux_up"], q["flux_down"]
    res["heating_rate"] = grab(g.graphe_heating_rate, ["heating_rate"], I, mu, z, L, M,
                               out["idx_up"], out["idx_down"], F0, mu0, tau, grd_alb, aer)["heating_rate"]
    return res


# --------------------------------------------------------------------------
def stage_small():
    ref = rh.load_reference()
    R = ref["I1_In"]
    # ---- phase matrices (small, all analytic families) ----
    ph = {}
    for M in (21, 41):
        mu = mu_grid(M)
        for name, g in (("iso", 0.0), ("rayleigh", 0.0), ("hg", 0.5), ("hg", 0.75), ("fwc", 0.0)):
            for mu0 in (0.5, 0.8):
                if name == "iso" and mu0 != 0.5:
                    continue
                P0, P = rh.phase_matrices(name, M, mu, mu0, g)
                ph[f"{name}_M{M}_mu0{mu0}_g{g}_P0"] = P0
                ph[f"{name}_M{M}_mu0{mu0}_g{g}_P"] = P
    np.savez_compressed(os.path.join(HERE, "phase_small.npz"), **ph)

    # ---- single-layer functions: several tau* regimes (extrapolation widths) ----
    sl = {}
    cases = [(80, 101, 0.05, 0.5, 0.9), (80, 101, 0.5, 0.5, 1.0), (60, 251, 2.0, 0.3, 0.9),
             (120, 101, 8.0, 0.5, 0.9), (60, 501, 0.3, 0.5, 0.95), (40, 201, 0.05, 0.7, 1.0),
             (70, 1201, 0.052, 0.5, 1.0)]
    for ci, (L, M, ts, mu0, alb) in enumerate(cases):
        mu = mu_grid(M)
        tau = np.linspace(0, ts, L)
        if M <= 101:
            P0, P = rh.phase_matrices("hg", M, mu, mu0, 0.5)
            pname = "hg"
        else:
            P0, P = np.ones(2 * M), 2 * np.ones((2 * M, 2 * M))
            pname = "iso"
        I1 = R.I1_NumInt(tau, mu, ts, mu0, P0, alb, M)
        J2 = R.Jn_NumInt(2, I1, tau, mu, ts, mu0, P, alb, M)
        I2 = R.In_NumInt(2, J2, I1, tau, mu, ts, mu0, P, alb, M, 0, 0)
        Js = smooth_source(tau, mu, ts)
        Is = R.In_NumInt(2, Js, I1, tau, mu, ts, mu0, P, alb, M, 0, 0)
        sl[f"c{ci}_params"] = np.array([L, M, ts, mu0, alb])
        sl[f"c{ci}_phase"] = np.array(pname)
        rows = np.arange(L) if M < 501 else np.unique(np.concatenate((np.arange(0, L, 6), [1, L - 2, L - 1])))
        sl[f"c{ci}_rows"] = rows
        sl[f"c{ci}_I1"], sl[f"c{ci}_J2"], sl[f"c{ci}_I2"], sl[f"c{ci}_Is"] = I1[rows], J2[rows], I2[rows], Is[rows]
        sl[f"c{ci}_mu12"] = np.array(R.mu_approx_In(mu, M))
    sl["ncases"] = np.array(len(cases))
    np.savez_compressed(os.path.join(HERE, "single_layer.npz"), **sl)

    # ---- three-region drivers on small grids ----
    dr = {}
    M = 41
    mu = mu_grid(M)
    P0a, Pa = rh.phase_matrices("rayleigh", M, mu, 0.5)
    runs = [
        ("thin", dict(nb_layers=60, tauStar_atm=0.124, grd_alb=0.15, alb_aer=0.97), 0.5),
        ("thick", dict(nb_layers=70, tauStar_atm=0.5, tauStar_aer=1.5, grd_alb=0.3, alb_aer=0.9, mu0=0.8), 0.8),
        ("mu0hit", dict(nb_layers=50, tauStar_atm=0.02, tauStar_aer=0.01, grd_alb=1, mu0=0.5), 0.5),
    ]
    for kind in ("specular", "lambertian"):
        for tag, kw, mu0 in runs:
            P0a, Pa = rh.phase_matrices("rayleigh", M, mu, mu0)
            P0h, Ph = rh.phase_matrices("hg", M, mu, mu0, 0.5)
            out = rh.run_driver(kind, phase={"atm": (P0a, Pa), "aer": (P0h, Ph)}, nb_angles=M,
                                aer_phase_fun="hg", **kw)
            key = f"{kind}_{tag}"
            dr[key + "_I"] = out["I"]
            no = len(out["I_saved"])
            keep = np.array([i for i in range(no) if i < 6 or i % 10 == 0 or i == no - 1])
            dr[key + "_order_ids"] = keep
            dr[key + "_orders"] = np.stack([out["I_saved"][i] for i in keep])
            dr[key + "_n"] = np.array(out["n"])
            dr[key + "_tau"] = out["tau"]
            dr[key + "_idx"] = np.array([out["idx_up"], out["idx_down"]])
            q = capture_quadratures(out["I"], out, M, mu0, kw.get("grd_alb"))
            for k, v in q.items():
                dr[f"{key}_{k}"] = v
            dr[key + "_kw"] = np.array(repr(dict(kw, nb_angles=M)))
    np.savez_compressed(os.path.join(HERE, "drivers_small.npz"), **dr)
    print("small done")


def _driver_record(out, M, mu0, grd_alb, row_stride, store_orders_rows=True):
    L = out["I"].shape[0]
    iu, idn = int(out["idx_up"]), int(out["idx_down"])
    rows = sorted(set([0, iu - 1, iu, idn, idn + 1, L - 1]))
    rec = {
        "n": np.array(out["n"]),
        "idx": np.array([iu, idn]),
        "tau": out["tau"],
        "rows": np.array(rows),
        "I_sub": out["I"][::row_stride].copy(),
        "I_rows": out["I"][rows].copy(),
        "order_max": np.array([np.max(np.abs(x)) for x in out["I_saved"]]),
        "order_sum": np.array([np.sum(x) for x in out["I_saved"]]),
    }
    if store_orders_rows:
        rec["order_rows"] = np.stack([x[rows] for x in out["I_saved"]])
    q = capture_quadratures(out["I"], out, M, mu0, grd_alb)
    rec.update(q)
    return rec


def stage_n1002():
    """M=501 (windowed columns, real extrapolation widths) at reduced L."""
    M = 501
    mu0 = 0.5
    P0a, Pa = ref_phase("rayleigh", M, mu0)
    P0h, Ph = ref_phase("hg", M, mu0, 0.5)
    ph = {"atm": (P0a, Pa), "aer": (P0h, Ph)}
    rec = {"phase_sub_atm": Pa[::25, ::25].copy(), "phase_sub_aer": Ph[::25, ::25].copy(),
           "P0_atm": P0a, "P0_aer": P0h,
           "colint_atm": np.array([np.trapz(Pa[:, n], mu_grid(M)) for n in range(0, 2 * M, 50)])}
    runs = [
        # EVA-like: idx = 10 in all regions
        ("eva_spec", "specular", dict(nb_layers=120, tauStar_atm=0.124, grd_alb=0.15, alb_aer=0.97)),
        ("eva_lamb", "lambertian", dict(nb_layers=120, tauStar_atm=0.124, grd_alb=0.15, alb_aer=0.97)),
        # very thin: idx = 2 everywhere -> windowed columns survive and feed the 2-point line
        ("thin_spec", "specular", dict(nb_layers=100, tauStar_atm=0.03, tauStar_aer=0.02, grd_alb=0.5)),
        # mixed widths: region 1 thin (idx=2), regions 2/3 idx=10
        ("mixed_spec", "specular", dict(nb_layers=140, tauStar_atm=0.05, tauStar_aer=0.4, grd_alb=0.15, alb_aer=0.9)),
        # thicker: idx = 20
        ("tau2_lamb", "lambertian", dict(nb_layers=160, tauStar_atm=0.3, tauStar_aer=1.4, grd_alb=0.3, alb_aer=0.85)),
    ]
    for tag, kind, kw in runs:
        t0 = time.time()
        out = rh.run_driver(kind, phase=ph, aer_phase_fun="hg", **kw)
        r = _driver_record(out, M, mu0, kw["grd_alb"], row_stride=10)
        for k, v in r.items():
            rec[f"{tag}_{k}"] = v
        rec[f"{tag}_kw"] = np.array(repr(dict(kw, kind=kind)))
        print(tag, "n =", out["n"], f"{time.time() - t0:.1f}s", flush=True)
    rec["tags"] = np.array([r[0] for r in runs])
    np.savez_compressed(os.path.join(HERE, "drivers_n1002.npz"), **rec)


def stage_default(which=None):
    """The 800 x 1002 scenarios (HG g=0.5 stands in for the log-normal Mie aerosol)."""
    M = 501
    mu0 = 0.5
    P0a, Pa = ref_phase("rayleigh", M, mu0)
    P0h, Ph = ref_phase("hg", M, mu0, 0.5)
    ph = {"atm": (P0a, Pa), "aer": (P0h, Ph)}
    eva = dict(tauStar_atm=0.124, grd_alb=0.15, alb_aer=0.97)
    wild = dict(tauStar_atm=0.124, tauStar_aer=0.0075, z_up=15, z_down=14, grd_alb=0.15, alb_aer=0.97)
    runs = {
        "eva_spec": ("specular", eva),
        "eva_lamb": ("lambertian", eva),
        "wildfire_lamb": ("lambertian", wild),
        "shipped_spec": ("specular", {}),  # literals exactly as shipped (rho=1, omega=1)
    }
    for tag, (kind, kw) in runs.items():
        if which and tag not in which:
            continue
        t0 = time.time()
        out = rh.run_driver(kind, phase=ph, aer_phase_fun="hg", **kw)
        r = _driver_record(out, M, mu0, kw.get("grd_alb", 1), row_stride=40)
        r["kw"] = np.array(repr(dict(kw, kind=kind)))
        np.savez_compressed(os.path.join(HERE, f"default_{tag}.npz"), **r)
        print(tag, "n =", out["n"], f"{time.time() - t0:.1f}s", flush=True)


def stage_thick():
    """Config-4 stand-in: thick single homogeneous FWC layer on a reduced grid."""
    ref = rh.load_reference()
    R = ref["I1_In"]
    L, M, ts, mu0, alb = 1000, 101, 30.0, 0.5, 0.9
    mu = mu_grid(M)
    tau = np.linspace(0, ts, L)
    P0, P = ref_phase("fwc", M, mu0)
    I1 = R.I1_NumInt(tau, mu, ts, mu0, P0, alb, M)
    I = I1.copy()
    In_1 = I1
    In = np.ones_like(I1)
    n = 1
    keep = {}
    ratios = []
    t0 = time.time()
    while max(max(In[0, M:] / I[0, M:]), max(In[L - 1, :M] / I[L - 1, :M])) >= 1e-4:
        ratios.append(max(max(In[0, M:] / I[0, M:]), max(In[L - 1, :M] / I[L - 1, :M])))
        n += 1
        J = R.Jn_NumInt(n, In_1, tau, mu, ts, mu0, P, alb, M)
        In = R.In_NumInt(n, J, In_1, tau, mu, ts, mu0, P, alb, M, 0, 0)
        In_1 = In
        I = I + In
        if n in (2, 3, 10, 50):
            keep[f"order{n}_sub"] = In[::50].copy()
        if n % 10 == 0:
            print("thick order", n, ratios[-1], f"{time.time() - t0:.0f}s", flush=True)
    np.savez_compressed(os.path.join(HERE, "thick_fwc.npz"), params=np.array([L, M, ts, mu0, alb]),
                        n=np.array(n), ratios=np.array(ratios), I_sub=I[::25].copy(),
                        I_toa=I[0].copy(), I_surf=I[L - 1].copy(), P0=P0, P_sub=P[::8, ::8].copy(), **keep)
    print("thick done n =", n)


def stage_forcing():
    """SOS_Aer_radiative_forcing / SOS_Aer_critical_albedo (SOS_Aer_critical_albedo.py:20-410)."""
    ref = rh.load_reference()
    tp = ref["tau_profile"].tau_profile
    import contextlib, io
    rec = {}
    cases = [("a", dict(L=60, M=41, mu0=0.5, ta=0.124, te=0.12, alb_aer=0.97, rho=0.15, z_up=25, z_down=17)),
             ("b", dict(L=80, M=41, mu0=0.8, ta=0.2, te=0.4, alb_aer=0.8, rho=0.3, z_up=30, z_down=10))]
    for tag, c in cases:
        L, M, mu0 = c["L"], c["M"], c["mu0"]
        mu = mu_grid(M)
        P0a, Pa = rh.phase_matrices("rayleigh", M, mu, mu0)
        P0h, Ph = rh.phase_matrices("hg", M, mu, mu0, 0.5)
        with contextlib.redirect_stdout(io.StringIO()):
            tau = tp(c["ta"], c["te"], 120, c["z_up"], c["z_down"], L)
        z = np.linspace(120, 0, L)
        iu, idn = int(np.argmin(np.abs(z - c["z_up"]))), int(np.argmin(np.abs(z - c["z_down"])))
        dtau_aer = c["te"] / (idn + 1 - iu)
        dtau_atm = c["ta"] / L
        F0 = np.pi / mu0
        forcing, critical = rh.critical_albedo_functions(c["ta"] + c["te"])
        args = (dtau_aer, c["ta"], dtau_atm, Ph, P0h, c["alb_aer"], Pa, P0a, 1.0, c["rho"], F0, mu, mu0, M, tau, L, iu, idn)
        rec[tag + "_toa_net_flux"] = np.array(forcing(0, *args))          # tauStar_aer == 0 -> returns the TOA net flux
        rec[tag + "_forcing"] = np.array(forcing(c["te"], *args))          # Q19: identically 0
        rec[tag + "_critical"] = np.array(critical(c["te"], dtau_aer, c["ta"], dtau_atm, Ph, P0h, Pa, P0a, 1.0, c["rho"],
                                                   F0, mu, mu0, M, tau, L, iu, idn))
        rec[tag + "_case"] = np.array(repr(c))
        print(tag, rec[tag + "_toa_net_flux"], rec[tag + "_forcing"], rec[tag + "_critical"], flush=True)
    np.savez_compressed(os.path.join(HERE, "forcing.npz"), **rec)


def stage_fwc3():
    """The FWC cloud as the aerosol of the three-region specular driver (what bench.py's workload does with a third of
    its scenarios), at the corners of the sweep: mu0 = 0.1 / 1.0 (1.0 sits ON the mu grid), omega_aer = 0.7 / 1.0,
    tau_aer = 0.5 / 0.0075; M = 251 (windowed columns exist: |mu| = 0.004, 0.008), reduced L."""
    M = 251
    rec = {}
    runs = [
        ("mu01", dict(nb_layers=100, mu0=0.1, tauStar_atm=0.124, tauStar_aer=0.5, grd_alb=0.05, alb_aer=0.7)),
        ("mu1", dict(nb_layers=100, mu0=1.0, tauStar_atm=0.124, tauStar_aer=0.0075, grd_alb=0.3, alb_aer=1.0)),
        ("mid", dict(nb_layers=120, mu0=0.6, tauStar_atm=0.124, tauStar_aer=0.12, grd_alb=0.15, alb_aer=0.9)),
    ]
    for tag, kw in runs:
        t0 = time.time()
        mu0 = kw["mu0"]
        P0a, Pa = ref_phase("rayleigh", M, mu0)
        P0f, Pf = ref_phase("fwc", M, mu0)
        out = rh.run_driver("specular", phase={"atm": (P0a, Pa), "aer": (P0f, Pf)}, nb_angles=M, aer_phase_fun="fwc", **kw)
        r = _driver_record(out, M, mu0, kw["grd_alb"], row_stride=5)
        for k, v in r.items():
            rec[f"{tag}_{k}"] = v
        rec[f"{tag}_kw"] = np.array(repr(dict(kw, nb_angles=M)))
        rec[f"{tag}_P0_aer"] = P0f
        print(tag, "n =", out["n"], f"{time.time() - t0:.1f}s", flush=True)
    rec["tags"] = np.array([r[0] for r in runs])
    np.savez_compressed(os.path.join(HERE, "drivers_fwc3.npz"), **rec)


def stage_fwc_table():
    fw = rh.load_reference()["fwc_data"]
    out = os.path.join(ROOT, "sos-radiative-transfer_b200", "data")
    os.makedirs(out, exist_ok=True)
    np.savez_compressed(os.path.join(out, "fwc_table.npz"), mu_fwc=np.asarray(fw.mu_fwc, dtype=np.float64),
                        phase_func_FWC=np.asarray(fw.phase_func_FWC, dtype=np.float64))
    print("fwc table", len(fw.mu_fwc))


if __name__ == "__main__":
    stage = sys.argv[1]
    if stage == "small":
        stage_small()
    elif stage == "n1002":
        stage_n1002()
    elif stage == "default":
        stage_default(sys.argv[2:] or None)
    elif stage == "thick":
        stage_thick()
    elif stage == "forcing":
        stage_forcing()
    elif stage == "fwc_table":
        stage_fwc_table()
    elif stage == "fwc3":
        stage_fwc3()
    else:
        raise SystemExit(f"unknown stage {stage}")
