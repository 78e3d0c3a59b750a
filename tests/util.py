"""Shared helpers for the parity tests: golden-case loading and record comparison."""
import glob
import os

import numpy as np

from geoac_b200 import abi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")
TOY = os.path.join(GOLD, "ToyAtmo.met")
PI = 3.141592653589793238462643

# FP64 relative tolerance of the CUDA path against oracle/reference for continuous outputs (north_star: 1e-9 away
# from caustics).  Amplitude-like quantities of rays with a tiny Jacobian are LISTED, not loosened.
RTOL = 1e-9


def golden_cases(prefix=""):
    names = sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLD, prefix + "*.npz")))
    return [n for n in names if not n.startswith("grid_")]


def is_rngdep(variant):
    return variant in (abi.GEOAC_3D_RNGDEP, abi.GEOAC_GLOBAL_RNGDEP)


def load_grid(d):
    """The synthetic range-dependent grid a golden case was traced on: ax0, ax1, axz, T, u, v, rho (as the loader returns them)."""
    name = str(d["grid"])
    if name.startswith("synth:"):
        # a full-size grid regenerated in memory, bit-identical to what the reference's loader reads from the node files the golden
        # vector was traced on (geoac_b200/synth.py: checked against load_met_grid on those 40 000 files when the vector was made)
        from geoac_b200 import synth
        return list(getattr(synth, name.split(":", 1)[1])())
    g = np.load(os.path.join(GOLD, name + ".npz"))
    return [g[k] for k in ("ax0", "ax1", "axz", "T", "u", "v", "rho")]


def profile_path(d, tmpdir=None):
    """The .met profile a stratified golden case was traced on: the shipped ToyAtmo.met, or the synthetic config-3 profile
    (geoac_b200/synth.py) written to a temporary file exactly as the golden generator wrote it."""
    if "profile" in d.files and str(d["profile"]) == "config3":
        import tempfile
        from geoac_b200 import synth
        path = os.path.join(tmpdir or tempfile.mkdtemp(prefix="g"), "c3.met")
        synth.write_met(path, synth.config3_profile())
        return path
    return TOY


def load_case(name):
    d = np.load(os.path.join(GOLD, name + ".npz"))
    kv = dict(s.split("=", 1) for s in d["keys"].tolist())
    return d, kv


def is_global(variant):
    return variant in (abi.GEOAC_GLOBAL, abi.GEOAC_GLOBAL_RNGDEP)


def apply_keys(variant, p, kv, vert_limit_default=None):
    """Map the reference mains' key=value options onto geoac_params exactly as the mains do."""
    p.bounces = int(kv.get("bounces", 2))
    p.calc_amp = int(kv.get("CalcAmp", 1))
    p.accum_per_segment = 1 if variant == abi.GEOAC_2D else int(kv.get("accum_mode", 0))
    p.freq = float(kv.get("freq", 0.1))
    p.z_grnd = float(kv.get("z_grnd", 0.0))
    p.tweak_abs = max(0.0, float(kv.get("abs_coeff", 0.3)))
    if "alt_max" in kv and not is_rngdep(variant):      # RngDep mains: SetPropRegion runs after parsing and overwrites it
        p.vert_limit = float(kv["alt_max"])
    if "rng_max" in kv and variant in (abi.GEOAC_2D, abi.GEOAC_3D, abi.GEOAC_GLOBAL):
        p.range_limit = float(kv["rng_max"])
    z_src = float(kv.get("z_src", 0.0))
    if variant in (abi.GEOAC_2D, abi.GEOAC_3D, abi.GEOAC_3D_RNGDEP):
        p.src[0] = float(kv.get("x_src", 0.0))
        p.src[1] = float(kv.get("y_src", 0.0))
        p.src[2] = z_src
    else:
        p.src[0] = z_src
        if "lat_src" in kv or variant == abi.GEOAC_GLOBAL:
            p.src[1] = float(kv.get("lat_src", 30.0)) * PI / 180.0
            p.src[2] = float(kv.get("lon_src", 0.0)) * PI / 180.0
    return p


def angles_rad(theta_deg, phi_deg):
    return np.asarray(theta_deg) * PI / 180.0, PI / 2.0 - np.asarray(phi_deg) * PI / 180.0


def compare_records(got, want, variant, calc_amp, rtol, label="", exact_discrete=True, amp_rtol=None):
    """Compare two record sets. Discrete outputs bit-exact; continuous outputs to `rtol` relative.
    Returns a list of human-readable problems (empty = pass) and a dict of max relative differences."""
    problems, stats = [], {}
    if exact_discrete:
        if not np.array_equal(got["status"], want["status"]):
            bad = np.argwhere(got["status"] != want["status"])
            problems.append(f"{label}: status differs at {len(bad)} slots, first {bad[:5].tolist()}")
        if not np.array_equal(got["n_steps"], want["n_steps"]):
            bad = np.argwhere(got["n_steps"] != want["n_steps"])
            problems.append(f"{label}: n_steps differs at {len(bad)} slots, first {bad[:5].tolist()}")
    m = (want["status"] == abi.ST_ARRIVAL) & (got["status"] == abi.ST_ARRIVAL)
    neq = abi.eq_count(variant, calc_amp)
    fields = list(range(neq)) + [abi.F_TRAVELTIME, abi.F_ATTEN, abi.F_TURNHEIGHT, abi.F_INCLINATION, abi.F_BACKAZ, abi.F_AUX]
    if calc_amp:
        fields.append(abi.F_AMPLITUDE)
    neq0 = abi.eq_count(variant, 0)
    alt_index = {abi.GEOAC_2D: 1, abi.GEOAC_3D: 2, abi.GEOAC_3D_RNGDEP: 2}.get(variant)     # Global: r ~ 6370 km
    for f in fields:
        a, b = got["rec"][f][m], want["rec"][f][m]
        if a.size == 0:
            continue
        # scale: |reference value| with a floor.  The arrival altitude z_k (first sub-ground point) is a sub-step residual
        # of order 1e-6 km left over from a path that climbed to the turning height, so -- like every position -- its error
        # is measured against the extent of the path (the turning height), not against the residual itself.
        if f == alt_index:
            scale = np.maximum(np.abs(b), np.maximum(want["rec"][abi.F_TURNHEIGHT][m], 1.0))
        elif f < 18:
            scale = np.maximum(np.abs(b), 1e-3)
        else:
            scale = np.maximum(np.abs(b), 1e-12 * max(1.0, float(np.max(np.abs(b)))))
        rel = np.abs(a - b) / scale
        stats[f] = float(rel.max())
        # auxiliary (launch-angle derivative) states and the amplitude built from their determinant amplify rounding
        # near caustics / grazing incidence: they may be given their own tolerance and are listed
        is_aux = (f == abi.F_AMPLITUDE) or (neq0 <= f < neq)
        tol = amp_rtol if (amp_rtol is not None and is_aux) else rtol
        if not np.all(rel <= tol):
            i = int(np.argmax(rel))
            problems.append(f"{label}: field {f} max rel diff {rel.max():.3e} > {tol:g} (got {a[i]!r}, want {b[i]!r}; "
                            f"{int((rel > tol).sum())} of {rel.size} records)")
    return problems, stats


def path_cases():
    return [n for n in golden_cases() if n.endswith("_path")]


def golden_paths(d, cap):
    """Raypath rows of a golden case (rows x {ray, state[0..2], amp, att, tt, bounce, step}) -> path [n][cap][PATH_NF], rows [n]."""
    n = len(d["theta_deg"])
    path = np.zeros((n, cap, abi.PATH_NF)); rows = np.zeros(n, dtype=np.int32)
    for r in d["path"]:
        i = int(r[0])
        assert rows[i] < cap
        path[i, rows[i]] = r[1:]
        rows[i] += 1
    return path, rows


def compare_paths(got_path, got_rows, want_path, want_rows, rtol, amp_rtol, label=""):
    """Row counts, bounce and step indices exact; positions / sums to rtol (positions relative to the path extent);
    the amplitude along the path to amp_rtol (it passes through caustics, where it is singular)."""
    problems = []
    if not np.array_equal(got_rows, want_rows):
        return [f"{label}: raypath row counts differ: {got_rows.tolist()} vs {want_rows.tolist()}"]
    for i, nr in enumerate(want_rows):
        g, w = got_path[i, :nr], want_path[i, :nr]
        if not (np.array_equal(g[:, 6], w[:, 6]) and np.array_equal(g[:, 7], w[:, 7])):
            problems.append(f"{label}: ray {i}: bounce/step indices differ")
            continue
        for f, tol in ((0, rtol), (1, rtol), (2, rtol), (4, rtol), (5, rtol), (3, amp_rtol)):
            scale = np.maximum(np.abs(w[:, f]), max(1.0 if f < 3 else 1e-300, 1e-9 * float(np.abs(w[:, f]).max()))) if f != 3 else np.abs(w[:, f]) + 1e-300
            rel = np.abs(g[:, f] - w[:, f]) / scale
            if not np.all(rel <= tol):
                j = int(np.argmax(rel))
                problems.append(f"{label}: ray {i} path field {f}: max rel diff {rel.max():.3e} > {tol:g} at row {j} (got {g[j, f]!r}, want {w[j, f]!r})")
    return problems


def golden_caustics(d, cap):
    """Caustic rows of a golden case (rows x {ray, state[0..2], tt, bounce, step}) -> caustic [n][cap][CAUSTIC_NF], rows [n]."""
    n = len(d["theta_deg"])
    caus = np.zeros((n, cap, abi.CAUSTIC_NF)); rows = np.zeros(n, dtype=np.int32)
    for r in d["caustic"]:
        i = int(r[0])
        assert rows[i] < cap
        caus[i, rows[i]] = r[1:]
        rows[i] += 1
    return caus, rows


def compare_caustics(got, got_rows, want, want_rows, rtol, label=""):
    """Event counts, bounce and step indices exact (the step at which the Jacobian changes sign is a discrete output);
    positions relative to the path extent and travel time to rtol."""
    if not np.array_equal(got_rows, want_rows):
        return [f"{label}: caustic counts differ: {got_rows.tolist()} vs {want_rows.tolist()}"]
    problems = []
    for i, nr in enumerate(want_rows):
        g, w = got[i, :nr], want[i, :nr]
        if not np.array_equal(g[:, 4:], w[:, 4:]):
            problems.append(f"{label}: ray {i}: caustic bounce/step indices differ: {g[:, 4:].tolist()} vs {w[:, 4:].tolist()}")
            continue
        rel = np.abs(g[:, :4] - w[:, :4]) / np.maximum(np.abs(w[:, :4]), 1.0)
        if nr and rel.max() > rtol:
            problems.append(f"{label}: ray {i}: caustic rows differ by {rel.max():.3e}")
    return problems


# ---- the CPU oracle over all host cores (fork: the children share the atmosphere tables and never touch CUDA) ----
_ORC = {}


def _orc_chunk(sl):
    po, variant, at, p, th, ph = _ORC["args"]
    return sl, po.trace(variant, at, p, th[sl], ph[sl])


def oracle_trace_parallel(po, variant, at, p, th, ph, chunk=8, max_procs=32):
    """po.trace split over the host cores; `at` must have been built before the call (shared copy-on-write)."""
    import multiprocessing as mp
    n = len(th)
    out = {"rec": np.zeros((abi.NFIELDS, n, p.bounces + 1)), "status": np.zeros((n, p.bounces + 1), dtype=np.int32),
           "n_steps": np.zeros((n, p.bounces + 1), dtype=np.int32)}
    if n == 0:
        return out
    _ORC["args"] = (po, variant, at, p, np.ascontiguousarray(th), np.ascontiguousarray(ph))
    slices = [slice(i, min(n, i + chunk)) for i in range(0, n, chunk)]
    nproc = max(1, min(len(os.sched_getaffinity(0)), max_procs, len(slices)))
    with mp.get_context("fork").Pool(nproc) as pool:
        for sl, o in pool.imap_unordered(_orc_chunk, slices):
            out["rec"][:, sl] = o["rec"]; out["status"][sl] = o["status"]; out["n_steps"][sl] = o["n_steps"]
    return out
