"""Near-threshold listing (geoac_b200/nearthreshold.py) on CPU: synthetic records around every branch threshold, and the
whole parity verdict on a real ray set with the g++ build of the device code (tests/host_emul, FMA-contracted like the GPU
build) standing in for the device -- against the oracle, exactly as tests/test_gpu_scale_parity.py does on the B200."""
import numpy as np

from geoac_b200 import abi, nearthreshold as nt
from tests import util


def _blank(n, n_rec):
    return {"rec": np.zeros((abi.NFIELDS, n, n_rec)), "status": np.zeros((n, n_rec), dtype=np.int32), "n_steps": np.zeros((n, n_rec), dtype=np.int32)}


def test_margin_flags_catch_every_threshold_and_taint_later_bounces():
    p = abi.GeoacParams()
    p.vert_limit, p.range_limit, p.calc_amp, p.bounces = 139.9, 10000.0, 1, 2
    out = _blank(8, 3)
    out["status"][:] = abi.ST_ARRIVAL
    out["n_steps"][:] = 9000
    out["rec"][abi.F_MARGIN][:] = -0.4
    out["rec"][abi.F_TURNHEIGHT][:] = 50.0
    out["rec"][abi.F_JACOBIAN][:] = 1e6
    for f in (4, 5, 6, 8, 9, 10):
        out["rec"][f][:] = 1e3
    out["rec"][abi.F_MARGIN][1, 0] = -3e-8                      # crossed the ground at the very start of the step
    out["rec"][abi.F_MARGIN][2, 1] = -1.0 + 2e-9                # ... at its very end
    out["status"][3, 1] = abi.ST_BREAK; out["status"][3, 2] = abi.ST_NONE
    out["rec"][abi.F_MARGIN][3, 1] = 4e-7                       # barely beyond the region limit
    out["rec"][abi.F_TURNHEIGHT][4, 0] = 139.9 - 1e-8           # turning point at the ceiling
    out["n_steps"][5, 2] = 2                                    # intercept would read solution[k-2]
    out["rec"][abi.F_JACOBIAN][6, 1] = 1e-3                     # |D| against terms of 1e6: caustic
    flag, reasons = nt.margin_flags(out, abi.GEOAC_3D, p)
    want = np.zeros((8, 3), dtype=bool)
    want[1, :] = True; want[2, 1:] = True; want[3, 1:] = True; want[4, :] = True; want[5, 2] = True; want[6, 1:] = True
    assert np.array_equal(flag, want), flag
    assert "ground crossing" in reasons[(1, 0)][0] and "region limit" in reasons[(3, 1)][0] and "ceiling" in reasons[(4, 0)][0]
    assert "caustic" in reasons[(6, 1)][0] and "follows a flagged bounce" in reasons[(1, 2)][0]
    lines = nt.listing(out, abi.GEOAC_3D, p, np.arange(8.0), np.zeros(8))
    assert len(lines) == int(want.sum()) and lines[0].startswith("ray 1 ")


def test_check_against_separates_problems_from_listed_entries():
    p = abi.GeoacParams(); p.vert_limit, p.range_limit, p.calc_amp, p.bounces = 139.9, 10000.0, 1, 0
    got = _blank(4, 1)
    got["status"][:] = abi.ST_ARRIVAL; got["n_steps"][:] = 100
    got["rec"][abi.F_MARGIN][:] = -0.5; got["rec"][abi.F_JACOBIAN][:] = 1e5
    for f in range(12):
        got["rec"][f][:] = 10.0 + f
    got["rec"][abi.F_TRAVELTIME][:] = 1000.0; got["rec"][abi.F_AMPLITUDE][:] = 1e-5; got["rec"][abi.F_TURNHEIGHT][:] = 40.0
    want = {k: v.copy() for k, v in got.items()}
    want["rec"][abi.F_AMPLITUDE][1, 0] *= 1.0 + 3e-8            # ill-conditioned: explained by the perturbation response below
    want["rec"][abi.F_AMPLITUDE][2, 0] *= 1.0 + 3e-8            # NOT explained
    want["n_steps"][3, 0] = 101                                 # discrete flip on a slot the perturbation also flips
    pert = {k: v.copy() for k, v in got.items()}
    pert["rec"][abi.F_AMPLITUDE][1, 0] *= 1.0 + 1e-8
    pert["n_steps"][3, 0] = 101
    cond = nt.conditioning(lambda a, b: pert, np.zeros(4), np.zeros(4), got, abi.GEOAC_3D, 1)
    tainted, _ = nt.margin_flags(got, abi.GEOAC_3D, p)
    problems, listed, stats, n_disc = nt.check_against(got, want, abi.GEOAC_3D, 1, tainted, cond)
    assert len(problems) == 1 and "ray 2" in problems[0] and "amplitude" in problems[0]
    assert [(i, b, name) for i, b, name, *_ in listed] == [(1, 0, "amplitude")] and n_disc == 1
    want["rec"][0][0, 0] += 1e-6                                # a position error is never excused
    problems, _, _, _ = nt.check_against(got, want, abi.GEOAC_3D, 1, tainted, cond)
    assert any("field 0" in q for q in problems)


def test_parity_verdict_with_the_host_emulation_as_device(oracle, capsys):
    """3-D stratified, theta 1..60 deg x 4 azimuths, 2 bounces (the ray set of SURVEY App. F): FMA contraction moves the
    auxiliary states of near-caustic arrivals by up to ~1e-7 while every position-like output stays below 1e-11.  The verdict must
    pass with those entries listed (each within 10x its response to a 1e-10 rad change of the launch angle)."""
    from tests import emul
    variant = abi.GEOAC_3D
    theta_deg = np.tile(np.arange(1.0, 60.5, 1.0), 4)
    phi_deg = np.repeat(np.array([-90.0, -30.0, 30.0, 90.0]), 60)
    th, ph = util.angles_rad(theta_deg, phi_deg)
    arrs = oracle.load_met_1d(util.TOY)
    at = oracle.atmo1d(False, *arrs)
    p = oracle.default_params(variant, at)
    p.bounces, p.calc_amp = 2, 1
    trace = lambda a, b: emul.trace(variant, p, arrs, a, b)
    got = trace(th, ph)
    want = oracle.trace(variant, at, p, th, ph)
    cond = nt.conditioning(trace, th, ph, got, variant, 1)
    tainted, _ = nt.margin_flags(got, variant, p)
    problems, listed, stats, n_disc = nt.check_against(got, want, variant, 1, tainted, cond, label="3d")
    with capsys.disabled():
        print(f"\n[emul vs oracle] {len(listed)} listed entries, worst " + (f"{max(l[3] for l in listed):.2e}" if listed else "-")
              + "; plain fields max " + f"{max(v for k, v in stats.items() if k < 4 or k in (18, 19, 20, 22, 23)):.1e}")
    assert not problems, "\n".join(problems)
    assert n_disc == 0 and not tainted.any()
    m = want["status"] == abi.ST_ARRIVAL
    rel = np.abs(got["rec"][abi.F_AMPLITUDE][m] - want["rec"][abi.F_AMPLITUDE][m]) / np.abs(want["rec"][abi.F_AMPLITUDE][m])
    assert np.quantile(rel, 0.9) <= 1e-9 and np.median(rel) < 1e-10
    # the oracle's own Jacobian field is GeoAc_Jacobian(solution, k): same quantity as the device's record field
    D = np.abs(got["rec"][abi.F_JACOBIAN][m] - want["rec"][abi.F_JACOBIAN][m]) / np.abs(want["rec"][abi.F_JACOBIAN][m])
    assert np.quantile(D, 0.9) < 1e-8
