"""Near-threshold / ill-conditioned ray listing (north_star: "discrete outputs bit-exact except for rays within a stated
epsilon of a branch threshold, which are listed"; "continuous outputs ... 1e-9 away from caustics").

Host-side post-processing of the records geoac_trace returns -- no numerics of the path live here, and nothing in this
module touches the CPU oracle.  Two mechanical criteria:

1. `margin_flags`: from one trace.  The kernel exports, per slot, how far into its last RK4 step the ray crossed the
   ground (ARRIVAL: GEOAC_F_MARGIN in (-1, 0)) or the violated region limit (BREAK: GEOAC_F_MARGIN in (0, 1]); the branch
   predicates are the strict comparisons of GeoAc_BreakCheck / GeoAc_GroundCheck applied after every step
   (reference Code/GeoAc/GeoAc.Solver.cpp:57-64).  A margin within `eps` of 0, -1 or 1 means that a rounding-level change
   of the state moves the crossing into the neighbouring step (step count +-1), and a turning height or end range within
   `eps_limit` of a region limit means the BreakCheck itself is within rounding.  Everything after such a slot on the same
   ray inherits the flag (the reflection restarts from a different state).  Also flagged: segments with fewer than three
   states (the intercept reads solution[k-2]) and arrivals whose Jacobian determinant D (GEOAC_F_JACOBIAN)
   is small against the sum of its own terms (kappa = sum|terms| / |D|, cancellation in GeoAc_Jacobian).

2. `conditioning`: from a second trace with every launch angle moved by `delta` radians (default 1e-10 -- eight orders
   below the 0.05 deg spacing of the launch grids, i.e. physically the same ray).  It yields, per arrival, the relative
   response of the amplitude, of D and of the auxiliary (launch-angle derivative) states, and every slot whose status or
   step count changed.  The auxiliary system is the linearisation of the ray equations; near caustic-forming rays its
   condition number with respect to the launch angle reaches 1e9 (measured on config 2), so rounding-level differences
   between two correct implementations (FMA contraction alone) show up at 1e-7 there while the ray's
   position, travel time and attenuation still agree to 1e-12.  A record entry may differ from the reference by more than
   1e-9 only where this response says the quantity is indeterminate at that level; such entries are LISTED with |D|.
"""
import numpy as np

from . import abi

PI = 3.141592653589793238462643


def _aux_fields(variant, calc_amp):
    neq0, neq = abi.eq_count(variant, 0), abi.eq_count(variant, calc_amp)
    return list(range(neq0, neq))


def jacobian_kappa(rec, variant):
    """Upper bound of sum|terms| / |D| for GeoAc_Jacobian at every slot (direction cosines bounded by 1); inf where D = 0."""
    y = rec
    a = np.abs
    if variant == abi.GEOAC_2D:                       # D = r (r' Z - z' R), 2DStratified.cpp:291-300
        S = a(y[0]) * (a(y[4]) + a(y[3]))
    elif variant == abi.GEOAC_3D:                     # 3DStratified.cpp:425-427
        S = a(y[5] * y[10]) + a(y[9] * y[6]) + a(y[4]) * (a(y[10]) + a(y[9])) + a(y[8]) * (a(y[6]) + a(y[5]))
    elif variant == abi.GEOAC_3D_RNGDEP:              # 3DRngDep.cpp:547-565
        S = a(y[7] * y[14]) + a(y[13] * y[8]) + a(y[6]) * (a(y[14]) + a(y[13])) + a(y[12]) * (a(y[8]) + a(y[7]))
    else:                                             # Global.cpp:594-608: r^2 cos(lat) (dr_ds ... ) with dt_ds ~ 1/r, dp_ds ~ 1/(r sin(lat))
        r = np.where(y[0] > 0, y[0], 1.0)
        sl, cl = np.maximum(a(np.sin(y[1])), 1e-300), a(np.cos(y[1]))
        S = r * r * cl * (a(y[7] * y[14]) + a(y[13] * y[8]) + a(y[6]) * (a(y[14]) / r + a(y[13]) / (r * sl))
                          + a(y[12]) * (a(y[8]) / r + a(y[7]) / (r * sl)))
    D = a(rec[abi.F_JACOBIAN])
    with np.errstate(divide="ignore", invalid="ignore"):
        return np.where(D > 0, S / D, np.inf)


def margin_flags(out, variant, params, eps=1e-6, eps_limit=1e-9, kappa_max=1e6):
    """Slots whose discrete outcome is within rounding of a branch threshold.  Returns (flag [n, n_rec] bool, reasons: dict
    (ray, bounce) -> text).  `params` is the geoac_params the batch was traced with."""
    st, ns, rec = out["status"], out["n_steps"], out["rec"]
    n, n_rec = st.shape
    flag = np.zeros((n, n_rec), dtype=bool)
    reasons = {}

    def mark(mask, text_fn):
        for i, b in np.argwhere(mask):
            flag[i, b] = True
            reasons.setdefault((int(i), int(b)), []).append(text_fn(int(i), int(b)))

    arr, brk = st == abi.ST_ARRIVAL, st == abi.ST_BREAK
    mg = rec[abi.F_MARGIN]
    mark(arr & ((mg > -eps) | (mg < -1.0 + eps)), lambda i, b: f"ground crossing at fraction {mg[i, b]:+.3e} of the last step")
    mark(brk & ((mg < eps) | (mg > 1.0 - eps)), lambda i, b: f"region limit crossed at fraction {mg[i, b]:.3e} of the last step")
    mark((arr | brk) & (ns < 3), lambda i, b: f"segment of {ns[i, b]} steps (intercept reads solution[k-2], App. A-20)")
    # turning height against the ceiling (Global: vert_limit is an absolute radius)
    glob = variant in (abi.GEOAC_GLOBAL, abi.GEOAC_GLOBAL_RNGDEP)
    ceil = params.vert_limit - (6370.0 if glob else 0.0)
    zt = rec[abi.F_TURNHEIGHT]
    mark(arr & (np.abs(zt - ceil) <= eps_limit * max(1.0, abs(ceil))), lambda i, b: f"turning height {zt[i, b]:.9f} km at the ceiling {ceil:g}")
    if variant in (abi.GEOAC_2D, abi.GEOAC_3D):
        rng = np.abs(rec[0]) if variant == abi.GEOAC_2D else np.hypot(rec[0], rec[1])
        mark(arr & (np.abs(rng - params.range_limit) <= eps_limit * params.range_limit), lambda i, b: f"range {rng[i, b]:.6f} km at the range limit")
    if variant in (abi.GEOAC_3D_RNGDEP, abi.GEOAC_GLOBAL_RNGDEP):
        o = 0 if variant == abi.GEOAC_3D_RNGDEP else 1
        for ax in range(2):
            span = params.box_max[ax] - params.box_min[ax]
            near = np.minimum(np.abs(rec[o + ax] - params.box_min[ax]), np.abs(rec[o + ax] - params.box_max[ax])) <= eps_limit * span
            mark(arr & near, lambda i, b, ax=ax: f"arrival on the edge of the region (axis {ax})")
    if params.calc_amp:
        kap = jacobian_kappa(rec, variant)
        mark(arr & (kap > kappa_max), lambda i, b: f"|D| = {abs(rec[abi.F_JACOBIAN][i, b]):.3e} is {kap[i, b]:.1e} times smaller than its terms (caustic)")
    # a flagged slot taints the rest of its ray
    tainted = np.logical_or.accumulate(flag, axis=1)
    for i, b in np.argwhere(tainted & ~flag):
        reasons.setdefault((int(i), int(b)), []).append("follows a flagged bounce of the same ray")
    return tainted, reasons


def _scale(variant, f, b, want_rec, m):
    """Magnitude an entry of field f is measured against: positions against the path extent (the arrival altitude is a
    sub-step residual), eikonal components against the unit eikonal vector, everything else against itself."""
    alt_index = {abi.GEOAC_2D: 1, abi.GEOAC_3D: 2, abi.GEOAC_3D_RNGDEP: 2}.get(variant)
    eik = {abi.GEOAC_2D: (2,), abi.GEOAC_3D: (3,)}.get(variant, (3, 4, 5))
    if f == alt_index:
        return np.maximum(np.abs(b), np.maximum(want_rec[abi.F_TURNHEIGHT], 1.0))
    if f in eik:
        return np.maximum(np.abs(b), 1.0)
    if f < 18:
        return np.maximum(np.abs(b), 1e-3)
    if f == abi.F_INCLINATION or f == abi.F_BACKAZ:
        return np.maximum(np.abs(b), 1.0)                      # degrees
    if f in (abi.F_AMPLITUDE, abi.F_JACOBIAN):
        return np.maximum(np.abs(b), 1e-300)
    top = float(np.max(np.abs(b[m]))) if m.any() else 1.0
    return np.maximum(np.abs(b), 1e-12 * max(1.0, top))


def compared_fields(variant, calc_amp):
    neq0, neq = abi.eq_count(variant, 0), abi.eq_count(variant, calc_amp)
    plain = list(range(neq0)) + [abi.F_TRAVELTIME, abi.F_ATTEN, abi.F_TURNHEIGHT, abi.F_INCLINATION, abi.F_BACKAZ, abi.F_AUX]
    aux = (list(range(neq0, neq)) + [abi.F_AMPLITUDE]) if calc_amp else []
    return plain, aux


def conditioning(trace_fn, theta, phi, out, variant, calc_amp, delta=1e-10):
    """Trace the batch again with every launch angle moved by `delta` rad and measure each slot's response.
    trace_fn(theta, phi) -> records.  Returns dict:
      flips [n, n_rec] bool    status or step count changed (tainting the rest of the ray),
      resp  {field: [n, n_rec]} relative response of every compared field (same scales as check_against), 0 where the slot is
                                not an arrival in both runs; plus "amp", "jac", "aux" (max over the auxiliary states)."""
    pert = trace_fn(np.asarray(theta) + delta, np.asarray(phi) + delta)
    st = out["status"]
    flips = (pert["status"] != st) | (pert["n_steps"] != out["n_steps"])
    flips = np.logical_or.accumulate(flips, axis=1)
    both = (st == abi.ST_ARRIVAL) & (pert["status"] == abi.ST_ARRIVAL)
    plain, auxf = compared_fields(variant, calc_amp)
    resp = {}
    for f in sorted(set(plain + auxf + [abi.F_AMPLITUDE, abi.F_JACOBIAN])):
        a, b = pert["rec"][f], out["rec"][f]
        with np.errstate(divide="ignore", invalid="ignore"):
            r = np.abs(a - b) / _scale(variant, f, b, out["rec"], both)
        resp[f] = np.where(both, r, 0.0)
    res = {"flips": flips, "resp": resp, "amp": resp[abi.F_AMPLITUDE], "jac": resp[abi.F_JACOBIAN], "delta": delta}
    aux = np.zeros(st.shape)
    for f in auxf:
        if f != abi.F_AMPLITUDE:
            aux = np.maximum(aux, resp[f])
    res["aux"] = aux
    return res


def listing(out, variant, params, theta_deg, phi_deg, cond=None, eps=1e-6, sens_min=1e-10, limit=200):
    """Human-readable list: slots flagged by margin_flags, and (with a conditioning result) slots that flipped under the
    perturbation or whose amplitude / auxiliary response exceeds sens_min."""
    tainted, reasons = margin_flags(out, variant, params, eps=eps)
    lines = []
    for (i, b), why in sorted(reasons.items()):
        lines.append(f"ray {i} (theta {theta_deg[i]:.4f}, azimuth {phi_deg[i]:.4f}) bounce {b}: " + "; ".join(why))
    if cond is not None:
        D = out["rec"][abi.F_JACOBIAN]
        for i, b in np.argwhere(cond["flips"] & ~tainted):
            lines.append(f"ray {i} (theta {theta_deg[i]:.4f}, azimuth {phi_deg[i]:.4f}) bounce {b}: status / step count changes under a "
                         f"{cond['delta']:g} rad change of the launch angle")
        s = np.maximum(cond["amp"], cond["aux"])
        for i, b in np.argwhere(s > sens_min):
            lines.append(f"ray {i} (theta {theta_deg[i]:.4f}, azimuth {phi_deg[i]:.4f}) bounce {b}: ill-conditioned auxiliary state: "
                         f"amplitude moves {cond['amp'][i, b]:.2e}, auxiliary states {cond['aux'][i, b]:.2e} (relative) under a {cond['delta']:g} rad "
                         f"change of the launch angle; |D| = {abs(D[i, b]):.3e}")
    if len(lines) > limit:
        lines = lines[:limit] + [f"... {len(lines) - limit} more"]
    return lines


def check_against(got, want, variant, calc_amp, tainted, cond, rtol=1e-9, cond_factor=10.0, label=""):
    """Parity verdict of `got` against reference records `want` with the listing applied:
       * status / step counts must be equal on every slot that is neither margin-flagged nor flipped by the perturbation;
       * every compared field of every unflagged arrival must agree to max(rtol, cond_factor x its own response to the
         perturbation of the launch angle).  For positions, travel time, attenuation, turning height the response is ~1e-13,
         so they are held to rtol; auxiliary states, D and amplitude near caustic-forming rays, and the eikonal / inclination of
         grazing rays after several reflections, respond at 1e-8 ... 1e-6 and are then LISTED with |D|.
    Returns (problems, listed, stats, n_listed_discrete)."""
    problems, listed = [], []
    excl = tainted | (cond["flips"] if cond is not None else False)
    bad_st = (got["status"] != want["status"]) & ~excl
    bad_ns = (got["n_steps"] != want["n_steps"]) & ~excl
    for name, bad in (("status", bad_st), ("n_steps", bad_ns)):
        if bad.any():
            w = np.argwhere(bad)
            problems.append(f"{label}: {name} differs on {len(w)} unflagged slots, first {w[:5].tolist()}")
    n_listed_discrete = int((((got["status"] != want["status"]) | (got["n_steps"] != want["n_steps"])) & excl).sum())
    m = (want["status"] == abi.ST_ARRIVAL) & (got["status"] == abi.ST_ARRIVAL) & ~excl
    plain, auxf = compared_fields(variant, calc_amp)
    D = np.abs(got["rec"][abi.F_JACOBIAN])
    stats = {}
    for f in plain + auxf:
        a, b = got["rec"][f], want["rec"][f]
        rel = np.where(m, np.abs(a - b) / _scale(variant, f, b, want["rec"], m), 0.0)
        stats[f] = float(rel.max()) if rel.size else 0.0
        if cond is None:
            allow = np.full(rel.shape, rtol)
            r_f = np.zeros(rel.shape)
        else:
            # the auxiliary states of an arrival are one coupled linear system: if any of them (or D, or the amplitude built from
            # them) responds to the perturbation, the whole set is ill-conditioned there
            r_f = np.maximum(np.maximum(cond["aux"], cond["amp"]), cond["jac"]) if f in auxf else cond["resp"][f]
            allow = np.maximum(rtol, cond_factor * r_f)
        over = rel > allow
        name = "amplitude" if f == abi.F_AMPLITUDE else (f"aux state {f}" if f in auxf else f"field {f}")
        if over.any():
            i, bb = np.unravel_index(np.argmax(np.where(over, rel, 0.0)), rel.shape)
            problems.append(f"{label}: {name}: {int(over.sum())} arrivals differ by more than max({rtol:g}, {cond_factor:g} x perturbation response), "
                            f"worst {rel[i, bb]:.3e} at ray {i} bounce {bb} (response {r_f[i, bb]:.2e}, |D| {D[i, bb]:.3e})")
        for i, bb in np.argwhere((rel > rtol) & ~over):
            listed.append((int(i), int(bb), name, float(rel[i, bb]), float(r_f[i, bb]), float(D[i, bb])))
    return problems, listed, stats, n_listed_discrete
