"""Run the UNMODIFIED reference (/root/reference) in this container.

TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is imported by the product
package; this file is additionally only usable where /root/reference exists
(the build container) -- it is how the golden fixtures under tests/golden/
were produced (tests/golden/make_golden.py) and how oracle/sos_oracle.py was
pinned.  It never copies reference code: the reference modules are imported
from where they lie, and the drivers (which are import-time scripts with
literal parameters, SOS_Aer_main_specular.py:19-94,482) are exec'd from their
own source text after substituting literal parameter lines.

Shims:
  * matplotlib / miepython are absent -> stub modules on sys.path
  * `I1_In` and `SOS_Aer_vdh_extract` are imported by the drivers but are not
    in the tree (SOS_Aer_main_specular.py:6,8) -> aliases / stubs
  * driver results are locals of SOS_Aer() -> captured with sys.setprofile
  * phase matrices can be injected by wrapping SOS_Aer_phase_func.phase_func
"""
from __future__ import annotations

import contextlib
import io
import os
import re
import sys
import tempfile
import types

REFERENCE_DIR = os.environ.get("SOS_REFERENCE_DIR", "/root/reference")

_loaded = {}


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_DIR, "SOS_Aer_I1_In.py"))


def _make_stub_dir() -> str:
    d = tempfile.mkdtemp(prefix="sos_ref_stubs_")
    os.makedirs(os.path.join(d, "matplotlib"))
    with open(os.path.join(d, "matplotlib", "__init__.py"), "w") as f:
        f.write("")
    with open(os.path.join(d, "matplotlib", "pyplot.py"), "w") as f:
        f.write(
            "def __getattr__(name):\n"
            "    def _noop(*a, **k):\n"
            "        return None\n"
            "    return _noop\n"
        )
    with open(os.path.join(d, "miepython.py"), "w") as f:
        f.write(
            "def __getattr__(name):\n"
            "    def _absent(*a, **k):\n"
            "        raise RuntimeError('miepython is not installed (stub): ' + name)\n"
            "    return _absent\n"
        )
    return d


def load_reference():
    """Import the reference modules (once) and return them in a dict."""
    if _loaded:
        return _loaded
    if not reference_available():
        raise RuntimeError(f"reference not found at {REFERENCE_DIR}")
    stubs = _make_stub_dir()
    scratch = tempfile.mkdtemp(prefix="sos_ref_cwd_")
    sys.path[:0] = [stubs, REFERENCE_DIR]
    cwd = os.getcwd()
    os.chdir(scratch)  # global_va loads/saves its .npy cache in cwd at import
    try:
        import SOS_Aer_I1_In  # noqa
        sys.modules["I1_In"] = SOS_Aer_I1_In
        vdh = types.ModuleType("SOS_Aer_vdh_extract")
        vdh.vdh = lambda *a, **k: None
        vdh.In_up_down = lambda *a, **k: None
        sys.modules["SOS_Aer_vdh_extract"] = vdh
        with contextlib.redirect_stdout(io.StringIO()):
            import SOS_Aer_In_limit  # noqa
            import SOS_Aer_global_va  # noqa
            import SOS_Aer_phase_func  # noqa
            import SOS_Aer_tau_profile  # noqa
            import SOS_Aer_graphe  # noqa
            import SOS_Aer_fwc_data  # noqa
    finally:
        os.chdir(cwd)
    _loaded.update(
        I1_In=SOS_Aer_I1_In,
        In_limit=SOS_Aer_In_limit,
        global_va=SOS_Aer_global_va,
        phase_func=SOS_Aer_phase_func,
        tau_profile=SOS_Aer_tau_profile,
        graphe=SOS_Aer_graphe,
        fwc_data=SOS_Aer_fwc_data,
        scratch=scratch,
        real_phase_func=SOS_Aer_phase_func.phase_func,
    )
    return _loaded


# literal lines of SOS_Aer() that can be substituted (SOS_Aer_main_specular.py:23-94)
_LITERALS = {
    "mu0": r"^    mu0 = 0\.5$",
    "z0": r"^    z0 = 120 .*$",
    "z_up": r"^    z_up = 25 .*$",
    "z_down": r"^    z_down = 17 .*$",
    "nb_layers": r"^    nb_layers = 800$",
    "tauStar_atm": r"^    tauStar_atm = 0\.104 .*$",
    "tauStar_aer": r"^    tauStar_aer = 0\.120 .*$",
    "grd_alb": r"^    grd_alb = 1\s*$",
    "alb_atm": r"^    alb_atm = 1\.0$",
    "alb_aer": r"^    alb_aer = 1\.0$",
    "nb_angles": r"^    nb_angles = 501 .*$",
    "atm_phase_fun": r"^    atm_phase_fun = 'rayleigh' .*$",
    "aer_phase_fun": r"^    aer_phase_fun = 'eva' .*$",
    "g_atm": r"^    g_atm = 0\.5$",
    "g_aer": r"^    g_aer = 0\.5$",
}


def run_driver(kind="specular", phase=None, threshold=None, **overrides):
    """Exec a reference driver with substituted literals; return SOS_Aer() locals.

    kind      'specular'  -> SOS_Aer_main_specular.py as shipped
              'lambertian' -> SOS_Aer_main_lambertian.py with "repair A"
                  (lines 274-276, which raise ValueError, are
                  replaced by line 274 of the specular file)
    phase     optional dict {'atm': (P0, P), 'aer': (P0, P)} injected instead
              of calling the reference's (slow) phase_func builders
    overrides values for the literal lines listed in _LITERALS
    """
    ref = load_reference()
    fname = {"specular": "SOS_Aer_main_specular.py", "lambertian": "SOS_Aer_main_lambertian.py"}[kind]
    path = os.path.join(REFERENCE_DIR, fname)
    with open(path, encoding="utf-8") as f:
        src = f.read()
    if kind == "lambertian":
        with open(os.path.join(REFERENCE_DIR, "SOS_Aer_main_specular.py"), encoding="utf-8") as f:
            spec_lines = f.read().split("\n")
        lines = src.split("\n")
        assert "scatt_surface = np.zeros(nb_angles)" in lines[273], lines[273]
        lines[273:276] = [spec_lines[273]]
        src = "\n".join(lines)
    for key, val in overrides.items():
        pat = _LITERALS[key]
        src, nsub = re.subn(pat, f"    {key} = {val!r}", src, count=1, flags=re.M)
        assert nsub == 1, f"literal {key} not found in {fname}"
    if threshold is not None:
        src, nsub = re.subn(r">= 0\.0001:", f">= {threshold!r}:", src, count=1)
        assert nsub == 1

    pf_mod = ref["phase_func"]
    if phase is not None:
        def injected(mol, *a, **k):
            P0, P = phase[mol]
            return P0.copy(), P.copy()
        pf_mod.phase_func = injected
    captured = {}

    def prof(frame, event, arg):
        if event == "return" and frame.f_code.co_name == "SOS_Aer":
            captured.update(frame.f_locals)

    cwd = os.getcwd()
    os.chdir(ref["scratch"])
    try:
        with contextlib.redirect_stdout(io.StringIO()):
            sys.setprofile(prof)
            try:
                exec(compile(src, path, "exec"), {"__name__": "sos_ref_driver"})
            finally:
                sys.setprofile(None)
    finally:
        os.chdir(cwd)
        pf_mod.phase_func = ref["real_phase_func"]
    return captured


def phase_matrices(name, nb_angles, mu, mu0, g=0.5):
    """Call the reference's own builders directly (bypassing its cwd cache)."""
    ref = load_reference()
    pf = ref["phase_func"]
    with contextlib.redirect_stdout(io.StringIO()), contextlib.redirect_stderr(io.StringIO()):
        if name == "iso":
            return pf.isotropic(nb_angles, mu)
        if name == "hg":
            return pf.henyey_greenstein(nb_angles, mu, mu0, g)
        if name == "rayleigh":
            return pf.rayleigh(nb_angles, mu, mu0)
        if name == "fwc":
            return pf.fwc(nb_angles, mu, mu0)
    raise ValueError(name)


def critical_albedo_functions(tauStar_tot):
    """The two functions of SOS_Aer_critical_albedo.py (:20-410) without its module-level script.

    The file is a script that runs a whole sweep at import (:414-503); only the source above the
    "MAIN SCRIPT" banner is exec'd.  SOS_Aer_radiative_forcing reads `tauStar_tot` as a module global
    (:39), so it is planted in the namespace."""
    load_reference()
    path = os.path.join(REFERENCE_DIR, "SOS_Aer_critical_albedo.py")
    with open(path, encoding="utf-8") as f:
        lines = f.read().split("\n")
    cut = next(i for i, l in enumerate(lines) if "MAIN SCRIPT" in l) - 1
    src = "\n".join(lines[:cut])
    ns = {"__name__": "sos_ref_critical", "tauStar_tot": tauStar_tot}
    with contextlib.redirect_stdout(io.StringIO()):
        exec(compile(src, path, "exec"), ns)
    def quiet(fn):
        def call(*a, **k):
            with contextlib.redirect_stdout(io.StringIO()):
                return fn(*a, **k)
        return call
    return quiet(ns["SOS_Aer_radiative_forcing"]), quiet(ns["SOS_Aer_critical_albedo"])
