"""GPU measurement: config-2 style batch with raypath capture (WriteRays=True rows every 25 steps) through geoac_trace_paths."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench

_, _, _, th, ph = bench.workload_angles("config2")
sel = np.arange(len(th)).reshape(-1, 60)[::10].ravel()          # every 10th azimuth: 21 600 rays
th, ph = th[sel].copy(), ph[sel].copy()
tr, p = bench.setup_tracer("config2", 0)
p.accum_per_segment = 1
tr.params = p
for cap in (2000,):
    tr.trace_paths(th[:64], ph[:64], 25, cap)                    # warm-up
    t0 = time.time(); out = tr.trace_paths(th, ph, 25, cap); dt = time.time() - t0
    steps = int(out["n_steps"].sum()); rows = int(out["path_rows"].sum())
    print("paths: %d rays, %d steps, %d rows (max %d per ray, cap %d), %.2f s wall -> %.0f rays/s, %.3g steps/s, kernel %.1f ms" %
          (len(th), steps, rows, int(out["path_rows"].max()), cap, dt, len(th) / dt, steps / dt, tr.last_stats()[1]))
t0 = time.time(); o2 = tr.trace(th, ph); dt = time.time() - t0
print("plain (same accumulation mode): %.2f s wall, kernel %.1f ms" % (dt, tr.last_stats()[1]))
