"""Eigenray search timing on the GPU box: geoac_eigenray_search (batched, GPU) next to the unmodified
reference's one-ray-at-a-time search (oracle/_ref/ref_eig3d, one host core) on the same receivers.
ToyAtmo.met, source at the origin, bounce counts 0-2, receivers on two rings (250 and 400 km)."""
import json, os, subprocess, sys, tempfile, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import geoac_b200 as g
from geoac_b200 import abi

TOY = os.path.join(ROOT, "tests", "golden", "ToyAtmo.met")
n_rcvr = int(sys.argv[1]) if len(sys.argv) > 1 else 64
n_ref = int(sys.argv[2]) if len(sys.argv) > 2 else 4
az = np.arange(n_rcvr) * (360.0 / n_rcvr) + 7.0
rad = np.where(np.arange(n_rcvr) % 2 == 0, 250.0, 400.0)
rcv = np.stack([rad * np.sin(az * np.pi / 180.0), rad * np.cos(az * np.pi / 180.0)], axis=1)

tr = g.Tracer(abi.GEOAC_3D, 0)
tr.set_atmosphere_1d(*g.load_met_1d(TOY))
tr.eigenray_search(rcv[:1], bnc_min=0, bnc_max=2)            # warm-up (context, staging)
t = time.perf_counter(); rows1, st1 = tr.eigenray_search(rcv[:1], bnc_min=0, bnc_max=2); t_one = time.perf_counter() - t
t = time.perf_counter(); rows, st = tr.eigenray_search(rcv, bnc_min=0, bnc_max=2); t_all = time.perf_counter() - t

exe = os.path.join(ROOT, "oracle", "_ref", "ref_eig3d")
ref_s, ref_found, got_found = [], 0, 0
sel = list(range(0, n_rcvr, max(1, n_rcvr // n_ref)))[:n_ref]
if os.path.exists(exe):
    for i in sel:
        with tempfile.TemporaryDirectory() as td:
            out = os.path.join(td, "o.bin")
            t = time.perf_counter()
            subprocess.check_call([exe, out, td, TOY, f"x_rcvr={float(rcv[i, 0])!r}", f"y_rcvr={float(rcv[i, 1])!r}", "bnc_min=0", "bnc_max=2"],
                                  stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
            ref_s.append(time.perf_counter() - t)
            a = np.fromfile(out); r = a[2:].reshape(int(a[0]), 8)
            mine = rows[rows[:, 0] == i]
            assert len(mine) == len(r) and np.array_equal(mine[:, [1, 2, 6]], r[:, [0, 1, 5]]), i
            f = r[:, 5] == 1
            assert np.allclose(mine[f][:, [7, 8]], r[f][:, [6, 7]], rtol=0, atol=1e-7), i
            ref_found += int(f.sum()); got_found += int((mine[:, 6] == 1).sum())
print(json.dumps({
    "receivers": n_rcvr, "bounce_counts": [0, 2], "eigenrays_found": st["found"], "estimate_calls": len(rows),
    "gpu_all_receivers_s": round(t_all, 3), "gpu_batches": st["rounds"], "gpu_rays": st["rays"],
    "gpu_one_receiver_s": round(t_one, 3), "gpu_one_receiver_batches": st1["rounds"],
    "reference_sampled_receivers": sel, "reference_s_each": [round(x, 2) for x in ref_s],
    "reference_extrapolated_all_s": round(float(np.mean(ref_s)) * n_rcvr, 1) if ref_s else None,
    "sampled_receivers_agree": bool(ref_s) and ref_found == got_found, "host_cores_used_by_reference": 1}))
