"""GPU: trace config 5 in full once and save every ray's RK4 step counts / statuses (scheduling design data)."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from geoac_b200 import abi

_, th_deg, ph_deg, th, ph = bench.workload_angles("config5")
t0 = time.time(); tr, p = bench.setup_tracer("config5", 0); print("setup %.1f s" % (time.time() - t0), flush=True)
t0 = time.time(); out = tr.trace(th, ph); t1 = time.time()
ns = out["n_steps"].astype(np.int32); st = out["status"].astype(np.int8)
tot = ns.sum(axis=1)
print("full config 5: %.2f s, total steps %.5g, mean %.0f, max %d, occ %.3f" % (t1 - t0, tot.sum(), tot.mean(), tot.max(), tr.last_lane_occupancy()))
print("quantiles", np.quantile(tot, [0.5, 0.9, 0.99, 0.999, 0.9999, 0.99999]).astype(int).tolist())
for thr in (50000, 100000, 200000, 500000, 999999, 1500000, 2500000):
    print("rays >", thr, int((tot > thr).sum()), "steps in them %.4g" % tot[tot > thr].sum())
print("LIMIT statuses:", int((st == abi.ST_LIMIT).sum()), "BREAK:", int((st == abi.ST_BREAK).sum()), "ARRIVAL:", int((st == abi.ST_ARRIVAL).sum()))
np.savez_compressed("gpurun_out/r2a_config5_steps.npz", n_steps=ns, status=st)
