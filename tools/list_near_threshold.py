#!/usr/bin/env python
"""List the near-threshold / ill-conditioned rays of a workload (north_star: "bit-exact except for rays within a stated epsilon
of a branch threshold, which are listed").  Runs on a B200 (the product has no CPU path):

    python tools/list_near_threshold.py --workload config2 [--every 100] [--eps 1e-6] [--delta 1e-10] [--sens 1e-9]

Traces the (sub-sampled) launch grid twice -- as given, and with every launch angle moved by --delta radians -- and prints
  * every slot whose ground / region-limit crossing lies within --eps of a step boundary, whose turning height or range sits on
    a region limit, whose segment is too short for the quadratic intercept, or whose Jacobian determinant cancels (margin_flags);
  * every slot whose status or step count changes under the perturbation;
  * every arrival whose amplitude or auxiliary state moves by more than --sens (relative) under the perturbation, with |D|.
These are the rays for which a bit-exact / 1e-9 comparison with the reference is not meaningful (geoac_b200/nearthreshold.py)."""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench                                    # noqa: E402  (workload definitions)
from geoac_b200 import abi, nearthreshold as nt  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="config2", choices=sorted(bench.WORKLOADS))
    ap.add_argument("--every", type=int, default=100, help="trace every N-th ray of the launch grid")
    ap.add_argument("--eps", type=float, default=1e-6)
    ap.add_argument("--delta", type=float, default=1e-10)
    ap.add_argument("--sens", type=float, default=1e-7, help="list arrivals whose amplitude or auxiliary states move by more than this (relative) under the perturbation")
    ap.add_argument("--limit", type=int, default=200)
    a = ap.parse_args()
    variant = bench.WORKLOADS[a.workload][0]
    _, th_deg, ph_deg, th, ph = bench.workload_angles(a.workload)
    idx = np.arange(0, len(th), max(1, a.every))
    th_deg, ph_deg, th, ph = th_deg[idx], ph_deg[idx], th[idx].copy(), ph[idx].copy()
    tr, p = bench.setup_tracer(a.workload, 0)
    out = tr.trace(th, ph)
    cond = nt.conditioning(tr.trace, th, ph, out, variant, p.calc_amp, delta=a.delta)
    lines = nt.listing(out, variant, tr.params, th_deg, ph_deg, cond=cond, eps=a.eps, sens_min=a.sens, limit=a.limit)
    tainted, _ = nt.margin_flags(out, variant, tr.params, eps=a.eps)
    arr = out["status"] == abi.ST_ARRIVAL
    s = np.maximum(cond["amp"], cond["aux"])
    print(f"{a.workload}: {len(th)} rays, {int(arr.sum())} arrivals; {int(tainted.sum())} slots within {a.eps:g} of a branch threshold, "
          f"{int((cond['flips'] & ~tainted).sum())} more flip under a {a.delta:g} rad perturbation, "
          f"{int((s > a.sens).sum())} arrivals with amplitude / auxiliary response above {a.sens:g}")
    for ln in lines:
        print(ln)


if __name__ == "__main__":
    main()
