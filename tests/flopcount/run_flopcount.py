"""TEST INFRASTRUCTURE: derive the algorithmic FP64 operation count per RK4 step of every variant by running the
per-ray device code with an op-counting scalar (tests/flopcount/flopcount.cpp) -- SURVEY.md 8d (i).

    python tests/flopcount/run_flopcount.py            # prints one JSON line per variant, writes tests/flopcount/flops.json
"""
import ctypes as C
import json
import os
import subprocess
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from geoac_b200 import abi, synth          # noqa: E402
from oracle import pyoracle as po          # noqa: E402  (loader + default parameters only)
from tests import emul, util               # noqa: E402

dp = C.POINTER(C.c_double)


def build():
    so = os.path.join(HERE, "libflopcount.so")
    src = os.path.join(HERE, "flopcount.cpp")
    csrc = os.path.join(ROOT, "geoac_b200", "csrc")
    deps = [src, os.path.join(HERE, "counted.hpp")] + [os.path.join(csrc, f) for f in os.listdir(csrc)]
    if not os.path.exists(so) or any(os.path.getmtime(d) > os.path.getmtime(so) for d in deps):
        subprocess.check_call(["g++", "-O1", "-std=c++17", "-fPIC", "-shared", src, "-o", so])
    return C.CDLL(so)


def main():
    L = build()
    PI = util.PI
    out = {}
    # stratified variants on the shipped profile (configs 1, 2) and the config-3 profile
    toy = po.load_met_1d(util.TOY)
    cases = [(abi.GEOAC_2D, toy, np.arange(2.0, 45.0, 6.0), np.full(8, -90.0), {}),
             (abi.GEOAC_3D, toy, np.tile(np.arange(2.0, 60.0, 8.0), 2), np.repeat([0.0, 135.0], 8), {}),
             (abi.GEOAC_GLOBAL, po.load_met_1d(util.TOY, global_taper=True), np.tile(np.arange(2.0, 50.0, 8.0), 2), np.repeat([0.0, 135.0], 6), {"bounces": 3})]
    for variant, prof, th_deg, ph_deg, kv in cases:
        at = po.atmo1d(variant == abi.GEOAC_GLOBAL, *prof)
        p = util.apply_keys(variant, po.default_params(variant, at), kv)
        tab, n = emul.make_table(variant == abi.GEOAC_GLOBAL, *prof)
        th, ph = util.angles_rad(th_deg, ph_deg)
        th = np.ascontiguousarray(th); ph = np.ascontiguousarray(ph)
        sys.stdout.flush()
        L.flopcount_1d(variant, C.byref(p), n, tab.ctypes.data_as(dp), len(th), th.ctypes.data_as(dp), ph.ctypes.data_as(dp))
    # range-dependent variants on coarse versions of the config 4 / 5 grids
    for variant, grid, src in ((abi.GEOAC_3D_RNGDEP, synth.config4_grid(21, 21, 300), None),
                               (abi.GEOAC_GLOBAL_RNGDEP, synth.config5_grid(31, 61, 300), (0.0, 35.0 * PI / 180.0, 0.0))):
        arrs = [np.ascontiguousarray(a, dtype=np.float64) for a in grid]
        at = po.atmo3d(variant == abi.GEOAC_GLOBAL_RNGDEP, *arrs)
        p = po.default_params(variant, at)
        p.bounces = 1
        if src:
            p.src[0], p.src[1], p.src[2] = src
        th, ph = util.angles_rad(np.array([5.0, 15.0, 25.0, 35.0]), np.array([10.0, 100.0, 190.0, 280.0]))
        th = np.ascontiguousarray(th); ph = np.ascontiguousarray(ph)
        L.flopcount_3d(variant, C.byref(p), len(arrs[0]), len(arrs[1]), len(arrs[2]), *[a.ctypes.data_as(dp) for a in arrs], len(th),
                       th.ctypes.data_as(dp), ph.ctypes.data_as(dp))


if __name__ == "__main__":
    main()
