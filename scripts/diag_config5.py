"""Diagnostic (GPU): longest rays of the config-5 slice and the latency of tracing them alone, serial vs cooperative."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from geoac_b200 import abi

_, th_deg, ph_deg, th, ph = bench.workload_angles("config5")
n = 100000
th, ph, th_deg, ph_deg = th[:n].copy(), ph[:n].copy(), th_deg[:n], ph_deg[:n]
tr, p = bench.setup_tracer("config5", 0)
t0 = time.time(); out = tr.trace(th, ph); t1 = time.time()
tot = out["n_steps"].sum(axis=1)
print("slice: %.2f s, total steps %.4g, mean %.0f, max %d, occ %.3f" % (t1 - t0, tot.sum(), tot.mean(), tot.max(), tr.last_lane_occupancy()))
print("quantiles", np.quantile(tot, [0.5, 0.9, 0.99, 0.999, 0.9999]).astype(int).tolist())
top = np.argsort(-tot)[:12]
for i in top:
    print(int(i), round(float(th_deg[i]), 2), round(float(ph_deg[i]), 2), out["n_steps"][i].tolist(), out["status"][i].tolist())
print("rays > 100k steps:", int((tot > 100000).sum()), " > 200k:", int((tot > 200000).sum()), " LIMIT status:", int((out["status"] == abi.ST_LIMIT).sum()))
# latency of the 32 longest rays alone (one packet), serial vs cooperative
idx = np.sort(np.argsort(-tot)[:32])
for coop, lpt in (("0", "0"), ("1", "2")):
    os.environ["GEOAC_B200_COOP"] = coop; os.environ["GEOAC_B200_LPT"] = lpt
    tr2, _ = bench.setup_tracer("config5", 0)
    t0 = time.time(); o2 = tr2.trace(th[idx], ph[idx]); dt = time.time() - t0
    s2 = o2["n_steps"].sum(axis=1)
    print("coop=%s lpt=%s: 32 longest rays alone: %.2f s, max steps %d -> %.1f us per step of the longest ray" % (coop, lpt, dt, s2.max(), 1e6 * dt / s2.max()))
