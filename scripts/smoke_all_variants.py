"""Small end-to-end run of every kernel family in one process (all five variants, raypath / caustic capture, eigenray search,
device-built node tables), on small batches, with the claim order forced on (GEOAC_B200_LPT=2) so that the scout and the
sort kernels run too.  (compute-sanitizer is closed on this pool; this is the quick does-everything-launch check.)"""
import os, sys
os.environ["GEOAC_B200_LPT"] = "2"
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import geoac_b200 as g
from geoac_b200 import abi, synth
TOY = os.path.join(ROOT, "tests", "golden", "ToyAtmo.met")
for variant, glob in ((abi.GEOAC_2D, False), (abi.GEOAC_3D, False), (abi.GEOAC_GLOBAL, True)):
    tr = g.Tracer(variant, 0)
    tr.set_atmosphere_1d(*g.load_met_1d(TOY, global_taper=glob))
    p = tr.params; p.bounces = 1; tr.params = p
    _, _, th, ph = g.prop_angles(2.0, 40.0, 0.5, -90.0, 95.0, 45.0)
    out = tr.trace(th, ph)
    print(variant, len(th), "rays", int((out["status"] == abi.ST_ARRIVAL).sum()), "arrivals", tr.last_kernel_launches(), "launches")
    if variant == abi.GEOAC_3D:
        q = tr.params; q.accum_per_segment = 1; tr.params = q
        o2 = tr.trace_paths(th[:40], ph[:40], stride=25, cap=2400, caustic_cap=16)
        q.accum_per_segment = 0; tr.params = q
        print("  paths rows", int(o2["path_rows"].sum()), "caustics", int(o2["caustic_rows"].sum()))
        rows, st = tr.eigenray_search([(-250.0, 0.0)], bnc_min=0, bnc_max=0)
        print("  eigenrays", st)
for variant, grid in ((abi.GEOAC_3D_RNGDEP, synth.config4_grid(12, 12, 60)), (abi.GEOAC_GLOBAL_RNGDEP, synth.config5_grid(12, 14, 60))):
    tr = g.Tracer(variant, 0)
    tr.set_atmosphere_3d(*grid)
    p = tr.params; p.bounces = 0
    if variant == abi.GEOAC_GLOBAL_RNGDEP:
        p.src[0], p.src[1], p.src[2] = 0.0, 35.0 * np.pi / 180.0, 0.0
    tr.params = p
    _, _, th, ph = g.prop_angles(5.0, 30.0, 0.2, 10.0, 100.0, 45.0)
    out = tr.trace(th, ph)
    print(variant, len(th), "rays", int((out["status"] == abi.ST_ARRIVAL).sum()), "arrivals", tr.last_kernel_launches(), "launches")
print("all variants ran")
