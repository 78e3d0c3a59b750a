"""Atmosphere ingest timing on the GPU box: config-4 grid (200x200x300), node tables built on the device
vs on the host (GEOAC_B200_HOST_TABLES=1), tables compared bit for bit; node-file parse with the thread pool vs one thread
on a 60x60x300 subset of the same grid written as .met files."""
import json, os, sys, tempfile, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import geoac_b200 as g
from geoac_b200 import abi, synth

out = {}
grid = synth.config4_grid()
n0, n1, nz = len(grid[0]), len(grid[1]), len(grid[2])
res = {}
for mode in ("device", "host"):
    os.environ["GEOAC_B200_HOST_TABLES"] = "1" if mode == "host" else "0"
    tr = g.Tracer(abi.GEOAC_3D_RNGDEP, 0)
    tr.set_atmosphere_3d(*grid)            # first call pays context set-up
    t = time.perf_counter(); tr.set_atmosphere_3d(*grid); out[f"set_atmosphere_3d_{mode}_s"] = round(time.perf_counter() - t, 3)
    res[mode] = tr.grid_tables(n0, n1, nz)
    del tr
out["tables_bitwise_equal"] = bool(all(np.array_equal(a.view(np.uint64), b.view(np.uint64)) for a, b in zip(res["device"], res["host"])))
out["grid"] = [n0, n1, nz]
with tempfile.TemporaryDirectory() as td:
    xs = np.linspace(-500, 500, 60); ys = np.linspace(-500, 500, 60)
    pre, l0, l1 = synth.write_config4_files(td, xs, ys)
    for thr in ("1", "0"):
        if thr == "0": os.environ.pop("GEOAC_B200_LOAD_THREADS", None)
        else: os.environ["GEOAC_B200_LOAD_THREADS"] = thr
        t = time.perf_counter(); a = g.load_met_grid(pre, l0, l1); out["parse_3600_files_%s_s" % ("1thread" if thr == "1" else "pool")] = round(time.perf_counter() - t, 3)
out["host_cores"] = os.cpu_count()
print(json.dumps(out))
