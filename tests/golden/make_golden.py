"""Generate the golden vectors under tests/golden/ by running the UNMODIFIED reference (oracle/_ref/ref_<variant>, built
by `make -C oracle ref` from /root/reference/Code).  Run in the build container only; the .npz files are committed so
that the GPU box (which has no /root/reference) can check the oracle and the CUDA path against the reference itself.

    python tests/golden/make_golden.py [case ...]
"""
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from geoac_b200 import abi            # noqa: E402
from oracle import pyoracle as po     # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")
TOY = os.path.join(GOLD, "ToyAtmo.met")

# name -> (variant, profile args, reference-driver keys)
CASES = {
    # config 1 of BASELINE.json in full: GeoAc2D -prop ToyAtmo.met (90 rays, 2 bounces)
    "2d_config1": (abi.GEOAC_2D, [TOY], dict()),
    "2d_noamp": (abi.GEOAC_2D, [TOY], dict(theta_min=2, theta_max=44, theta_step=6, CalcAmp=0, bounces=1)),
    # 3-D stratified: a 1/100-style subsample of config 2's grid, both accumulation conventions, z_grnd/z_src/freq variants
    "3d_sub": (abi.GEOAC_3D, [TOY], dict(theta_min=1, theta_max=60.5, theta_step=4, phi_min=-90, phi_max=90, phi_step=45, bounces=2)),
    "3d_segmode": (abi.GEOAC_3D, [TOY], dict(theta_min=3, theta_max=45, theta_step=7, phi_min=-135, phi_max=180, phi_step=105, bounces=2, accum_mode=1)),
    "3d_noamp": (abi.GEOAC_3D, [TOY], dict(theta_min=2, theta_max=50, theta_step=8, phi_min=0, phi_max=300, phi_step=100, bounces=1, CalcAmp=0)),
    # spherical stratified (config 3's variant) on the shipped profile: default source lat 30 lon 0, both accumulation modes
    "global_sub": (abi.GEOAC_GLOBAL, [TOY], dict(theta_min=3, theta_max=48, theta_step=9, phi_min=-90, phi_max=135, phi_step=75, bounces=3)),
    "global_segmode": (abi.GEOAC_GLOBAL, [TOY], dict(theta_min=5, theta_max=35, theta_step=10, phi_min=20, phi_max=290, phi_step=135, bounces=1, accum_mode=1,
                                                       lat_src=-45, lon_src=170, rng_max=800, z_src=3.5, freq=0.7)),
    "global_noamp": (abi.GEOAC_GLOBAL, [TOY], dict(theta_min=4, theta_max=40, theta_step=12, phi_min=45, phi_max=225, phi_step=180, bounces=2, CalcAmp=0)),
    "3d_elevated": (abi.GEOAC_3D, [TOY], dict(theta_min=-10, theta_max=40, theta_step=10, azimuth=-60, bounces=2, z_src=12.5, z_grnd=1.2, freq=0.5, rng_max=600, alt_max=120)),
}


def main(names):
    for name in names:
        variant, prof, kv = CASES[name]
        with tempfile.TemporaryDirectory() as td:
            ref, info = po.run_ref(variant, prof, os.path.join(td, "o.bin"), **kv)
        out = os.path.join(GOLD, name + ".npz")
        np.savez_compressed(out, variant=variant, keys=np.array(sorted(f"{k}={v}" for k, v in kv.items())),
                            theta_deg=ref["theta_deg"], phi_deg=ref["phi_deg"], rec=ref["rec"], status=ref["status"],
                            n_steps=ref["n_steps"], eq_cnt=ref["eq_cnt"], vert_limit=info["vert_limit"])
        print(f"{name}: {len(ref['theta_deg'])} rays, {ref['total_steps']} steps, arrivals {(ref['status'] == 1).sum()}, "
              f"breaks {(ref['status'] == 2).sum()} -> {os.path.getsize(out)} bytes")


if __name__ == "__main__":
    main(sys.argv[1:] or list(CASES))
