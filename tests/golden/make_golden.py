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

from tests.golden import make_grid   # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")
TOY = os.path.join(GOLD, "ToyAtmo.met")

# synthetic range-dependent grids (tests/golden/make_grid.py); sources are placed OFF the node lines on purpose: on a
# node line the reference's cell choice depends on the spline cursor left behind by earlier look-ups
from geoac_b200 import synth   # noqa: E402


def _c3_profile(d):
    path = os.path.join(d, "c3.met")
    synth.write_met(path, synth.config3_profile())
    return [path]


GRIDS = {
    # coarse node grids carrying the SAME analytic atmospheres bench.py uses for configs 4 / 5 (geoac_b200/synth.py)
    "grid_c4": dict(is_global=False, build=lambda d: synth.write_config4_files(d, np.arange(-500.0, 501.0, 250.0), np.arange(-500.0, 501.0, 200.0), nz=300)),
    "grid_c5": dict(is_global=True, build=lambda d: synth.write_config5_files(d, np.arange(25.0, 46.0, 5.0), np.arange(-12.0, 13.0, 4.0), nz=300)),
    "grid_cart": dict(is_global=False, build=lambda d: make_grid.write_cartesian(d, np.arange(-500.0, 501.0, 200.0), np.arange(-450.0, 451.0, 150.0))),
    "grid_glob": dict(is_global=True, build=lambda d: make_grid.write_global(d, np.arange(20.0, 51.0, 6.0), np.arange(-15.0, 16.0, 5.0))),
}

# name -> (variant, profile args | grid name, reference-driver keys)
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
    # range-dependent variants on the synthetic grids above (config 4 / 5 style perturbations, tiny node counts)
    "3drngdep_sub": (abi.GEOAC_3D_RNGDEP, "grid_cart", dict(theta_min=5, theta_max=36, theta_step=10, phi_min=-60, phi_max=165, phi_step=105, bounces=1,
                                                            x_src=13.7, y_src=-21.3)),
    "3drngdep_noamp": (abi.GEOAC_3D_RNGDEP, "grid_cart", dict(theta_min=8, theta_max=30, theta_step=11, phi_min=30, phi_max=300, phi_step=135, bounces=1,
                                                              x_src=-120.4, y_src=88.8, CalcAmp=0, accum_mode=1, z_src=2.5)),
    "globalrngdep_sub": (abi.GEOAC_GLOBAL_RNGDEP, "grid_glob", dict(theta_min=5, theta_max=36, theta_step=10, phi_min=-60, phi_max=165, phi_step=105, bounces=1,
                                                                    lat_src=33.3, lon_src=1.7)),
    "globalrngdep_noamp": (abi.GEOAC_GLOBAL_RNGDEP, "grid_glob", dict(theta_min=8, theta_max=30, theta_step=11, phi_min=30, phi_max=300, phi_step=135, bounces=1,
                                                                      lat_src=36.1, lon_src=-3.4, CalcAmp=0, accum_mode=1, z_src=2.5)),
    # the synthetic atmospheres of BASELINE configs 3-5 themselves, on a few rays / a coarse node grid
    "global_c3": (abi.GEOAC_GLOBAL, _c3_profile, dict(theta_min=4, theta_max=40, theta_step=12, phi_min=30, phi_max=300, phi_step=90, bounces=5,
                                                     lat_src=30, lon_src=0, rng_max=3000)),
    "3drngdep_c4": (abi.GEOAC_3D_RNGDEP, "grid_c4", dict(theta_min=6, theta_max=46, theta_step=20, phi_min=20, phi_max=200, phi_step=140, bounces=2,
                                                         x_src=0, y_src=0)),
    "globalrngdep_c5": (abi.GEOAC_GLOBAL_RNGDEP, "grid_c5", dict(theta_min=6, theta_max=46, theta_step=20, phi_min=20, phi_max=200, phi_step=140, bounces=2,
                                                                 lat_src=35, lon_src=0)),
    # raypath rows (WriteRays=True: one row every 25 steps with the amplitude at that point), all five variants
    "2d_path": (abi.GEOAC_2D, [TOY], dict(theta_min=5, theta_max=36, theta_step=15, bounces=1, path_stride=25, caustics=1)),
    "3d_path": (abi.GEOAC_3D, [TOY], dict(theta_min=5, theta_max=46, theta_step=20, phi_min=30, phi_max=210, phi_step=170, bounces=1, accum_mode=1, path_stride=25, caustics=1)),
    "global_path": (abi.GEOAC_GLOBAL, [TOY], dict(theta_min=10, theta_max=31, theta_step=20, phi_min=60, phi_max=60, phi_step=1, bounces=1, accum_mode=1, path_stride=25, caustics=1)),
    "3drngdep_path": (abi.GEOAC_3D_RNGDEP, "grid_cart", dict(theta_min=12, theta_max=33, theta_step=20, phi_min=70, phi_max=70, phi_step=1, bounces=1,
                                                             x_src=13.7, y_src=-21.3, accum_mode=1, path_stride=25, caustics=1)),
    "globalrngdep_path": (abi.GEOAC_GLOBAL_RNGDEP, "grid_glob", dict(theta_min=12, theta_max=33, theta_step=20, phi_min=70, phi_max=70, phi_step=1, bounces=1,
                                                                     lat_src=33.3, lon_src=1.7, accum_mode=1, path_stride=25, caustics=1)),
    # -interactive of GeoAc3D plots one ray with a row every 10 steps (Code/GeoAc3D_main.cpp:409)
    "3d_interactive_path": (abi.GEOAC_3D, [TOY], dict(theta_min=40, theta_max=40, theta_step=1, phi_min=-45, phi_max=-45, phi_step=1, bounces=1, accum_mode=1, path_stride=10, caustics=1)),
    "3d_elevated": (abi.GEOAC_3D, [TOY], dict(theta_min=-10, theta_max=40, theta_step=10, azimuth=-60, bounces=2, z_src=12.5, z_grnd=1.2, freq=0.5, rng_max=600, alt_max=120)),
}


# The full-size config-4 golden (`3drngdep_c4full`: the unmodified reference on the 200 x 200 x 300 node grid of BASELINE config 4) is
# not made by main(): write the 40 000 node files with synth.write_config4_files("/tmp/c4", linspace(-500, 500, 200) twice, 300), check
# that oracle.load_met_grid on them equals synth.config4_grid_from_files() bit for bit (it does: all seven arrays), run
#   (ulimit -s unlimited; oracle/_ref/ref_3drngdep /tmp/c4full.bin /tmp/c4/p /tmp/c4/x.loc /tmp/c4/y.loc theta_min=6 theta_max=46
#    theta_step=10 phi_min=20 phi_max=200 phi_step=140 bounces=2 x_src=0 y_src=0)
# (24 s load, 12 s for the 10 rays) and store pyoracle.read_ref_bin("/tmp/c4full.bin") with grid = "synth:config4_grid_from_files".


def main(names):
    for name in names:
        variant, prof, kv = CASES[name]
        extra = {}
        with tempfile.TemporaryDirectory(dir="/tmp", prefix="g") as td:       # short paths: the reference's name buffers are char[50]
            if callable(prof):
                prof = prof(td)
                extra["profile"] = "config3"
            if isinstance(prof, str):
                grid = GRIDS[prof]
                files = list(grid["build"](td))
                gpath = os.path.join(GOLD, prof + ".npz")
                arrs = po.load_met_grid(*files, is_global=grid["is_global"])
                np.savez_compressed(gpath, **dict(zip(["ax0", "ax1", "axz", "T", "u", "v", "rho"], arrs)))
                extra["grid"] = prof
                prof = files
            ref, info = po.run_ref(variant, prof, os.path.join(td, "o.bin"), **kv)
        out = os.path.join(GOLD, name + ".npz")
        if "path" in ref:
            extra["path"] = ref["path"]
        if "caustic" in ref:
            extra["caustic"] = ref["caustic"]
        np.savez_compressed(out, variant=variant, **extra, keys=np.array(sorted(f"{k}={v}" for k, v in kv.items())),
                            theta_deg=ref["theta_deg"], phi_deg=ref["phi_deg"], rec=ref["rec"], status=ref["status"],
                            n_steps=ref["n_steps"], eq_cnt=ref["eq_cnt"], vert_limit=info["vert_limit"])
        print(f"{name}: {len(ref['theta_deg'])} rays, {ref['total_steps']} steps, arrivals {(ref['status'] == 1).sum()}, "
              f"breaks {(ref['status'] == 2).sum()} -> {os.path.getsize(out)} bytes")


if __name__ == "__main__":
    main(sys.argv[1:] or list(CASES))
