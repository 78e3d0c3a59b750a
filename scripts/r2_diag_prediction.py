"""GPU diagnosis: predicted (cost scout) vs actual RK4 step counts of one rank's share of the config-5 split."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
R, W = (int(x) for x in (sys.argv[1] if len(sys.argv) > 1 else "7/8").split("/"))
_, th_deg, ph_deg, th, ph = bench.workload_angles("config5")
mine = bench.shard_indices(len(th), R, W)
th, ph, th_deg, ph_deg = th[mine].copy(), ph[mine].copy(), th_deg[mine], ph_deg[mine]
tr, p = bench.setup_tracer("config5", 0)
t0 = time.time(); out = tr.trace(th, ph); dt = time.time() - t0
pred = tr.predicted_costs(len(th)).astype(np.int64)
act = out["n_steps"].sum(axis=1).astype(np.int64)
print("rank %d/%d: %.2f s, schedule %s" % (R, W, dt, tr.last_schedule()))
n = len(th) // 32 * 32
P, A = pred[:n].reshape(-1, 32).max(axis=1), act[:n].reshape(-1, 32).max(axis=1)
thr = pred.sum() / 37888
print("avg lane work (predicted) %.0f; packets predicted long %d, actually above it %d" % (thr, (P > thr).sum(), (A > thr).sum()))
bad = np.argsort(-(A - P))[:12]
for g in bad:
    i = g * 32 + int(np.argmax(act[g * 32:(g + 1) * 32]))
    print("packet %d: predicted max %d, actual max %d (ray theta %.2f az %.2f, segments %s)" % (g, P[g], A[g], th_deg[i], ph_deg[i], out["n_steps"][i].tolist()))
print("rays: corr(pred, act) = %.4f; median act/pred %.3f; worst underestimates:" % (np.corrcoef(pred, act)[0, 1], np.median(act / np.maximum(pred, 1))))
for i in np.argsort(-(act - pred))[:8]:
    print("   ray %d theta %.2f az %.2f predicted %d actual %d %s" % (i, th_deg[i], ph_deg[i], pred[i], act[i], out["status"][i].tolist()))
