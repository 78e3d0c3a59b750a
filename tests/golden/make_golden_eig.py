"""Golden vectors of the eigenray search from the UNMODIFIED reference: oracle/_ref/ref_eig3d[rngdep]
(oracle/ref_eig_driver.cpp linked against Code/GeoAc/GeoAc.Eigenray.cpp) -> tests/golden/eig/<case>.npz.
Run in the build container only (needs /root/reference):   python tests/golden/make_golden_eig.py [case ...]

rows [n][8]: { n_bnc, estimate_ok, theta_est, phi_est, theta_next, eigenray_found, theta_final, phi_final } per
GeoAc_EstimateEigenray call; `text` holds the reference's own <title>_results.dat (8 significant digits)."""
import os
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from geoac_b200 import abi            # noqa: E402
from tests.golden import make_golden as mg   # noqa: E402

OUT = os.path.join(mg.GOLD, "eig")

CASES = {
    # the shipped example: GeoAc3D -eig_search ToyAtmo.met (receiver 250 km west), 0 and 1 bounces
    "eig3d_axis": (abi.GEOAC_3D, [mg.TOY], dict(bnc_min=0, bnc_max=1)),
    # off-axis receivers (cross wind -> azimuth refinement passes), up to 2 bounces, elevated source
    "eig3d_cross": (abi.GEOAC_3D, [mg.TOY], dict(x_rcvr=-280, y_rcvr=130, bnc_min=0, bnc_max=2)),
    "eig3d_east": (abi.GEOAC_3D, [mg.TOY], dict(x_rcvr=310, y_rcvr=-75, bnc_min=0, bnc_max=1, z_src=1.5, azimuth_err_lim=0.5, theta_min=2, theta_max=40)),
    "eig3d_north": (abi.GEOAC_3D, [mg.TOY], dict(x_rcvr=40, y_rcvr=420, bnc_min=1, bnc_max=2, azimuth_err_lim=0.05)),
    # a tight azimuth limit forces the 4th / 5th correction passes, whose inclination step adapts ray by ray (Modify_d_theta)
    "eig3d_adaptive": (abi.GEOAC_3D, [mg.TOY], dict(x_rcvr=-280, y_rcvr=130, bnc_min=0, bnc_max=0, azimuth_err_lim=0.0004)),
    "eig3d_far": (abi.GEOAC_3D, [mg.TOY], dict(x_rcvr=-520, y_rcvr=60, bnc_min=1, bnc_max=2, azimuth_err_lim=0.4)),
    # range-dependent Cartesian variant on the synthetic grid of the parity cases
    "eig3drngdep": (abi.GEOAC_3D_RNGDEP, "grid_cart", dict(x_src=13.7, y_src=-21.3, x_rcvr=-230, y_rcvr=95, bnc_min=0, bnc_max=1)),
    # spherical variants (Code/GeoAc/GeoAc.Eigenray.Global.cpp): the shipped default (receiver 2.5 deg west) and an oblique one
    "eigglobal_w": (abi.GEOAC_GLOBAL, [mg.TOY], dict(bnc_min=0, bnc_max=1)),
    "eigglobal_nw": (abi.GEOAC_GLOBAL, [mg.TOY], dict(lat_src=41.5, lon_src=12.25, lat_rcvr=42.6, lon_rcvr=8.9, bnc_min=0, bnc_max=1, z_src=0.8, azimuth_err_lim=0.3)),
    "eigglobalrngdep": (abi.GEOAC_GLOBAL_RNGDEP, "grid_glob", dict(lat_src=33.3, lon_src=1.7, lat_rcvr=34.1, lon_rcvr=-1.2, bnc_min=0, bnc_max=0)),
    # -eig_direct: GeoAc_3DEigenray_LM alone from a user estimate (phi_est given as an azimuth, as on the command line).
    # NB the two GeoAc3D (stratified) cases are NOT reproducible bit for bit: in this mode the reference reads M_Comps
    # uninitialised (Eigenray.cpp:130-135 tests GeoAc_AtmoStrat before GeoAc_ConfigureCalcAmp :146 has set it), so its LM
    # iterates depend on stack garbage; the tests only require the same eigenray within the search tolerance there.
    "eigdirect3d": (abi.GEOAC_3D, [mg.TOY], dict(direct=1, theta_est=24.0, phi_est=-90.6, bounces=0)),
    "eigdirect3d_bnc": (abi.GEOAC_3D, [mg.TOY], dict(direct=1, x_rcvr=-520, y_rcvr=60, theta_est=25.5, phi_est=-83.3, bounces=1)),
    "eigdirect3drngdep": (abi.GEOAC_3D_RNGDEP, "grid_cart", dict(direct=1, x_src=13.7, y_src=-21.3, x_rcvr=-230, y_rcvr=95, theta_est=25.2, phi_est=-62.9, bounces=0)),
    "eigdirectglobal": (abi.GEOAC_GLOBAL, [mg.TOY], dict(direct=1, theta_est=7.0, phi_est=-89.0, bounces=0)),
}

EXE = {abi.GEOAC_3D: "ref_eig3d", abi.GEOAC_3D_RNGDEP: "ref_eig3drngdep", abi.GEOAC_GLOBAL: "ref_eigglobal", abi.GEOAC_GLOBAL_RNGDEP: "ref_eigglobalrngdep"}


def main(names):
    os.makedirs(OUT, exist_ok=True)
    for name in names:
        variant, prof, kv = CASES[name]
        extra = {}
        with tempfile.TemporaryDirectory(dir="/tmp", prefix="g") as td:
            if isinstance(prof, str):
                files = list(mg.GRIDS[prof]["build"](td))
                extra["grid"] = prof
                prof = files
            exe = os.path.join(ROOT, "oracle", "_ref", EXE[variant])
            wd = os.path.join(td, "w")
            os.makedirs(wd)
            out = os.path.join(td, "o.bin")
            subprocess.check_call([exe, out, wd] + list(prof) + [f"{k}={v}" for k, v in kv.items()], stdout=subprocess.DEVNULL)
            a = np.fromfile(out)
            rows = a[2:].reshape(int(a[0]), 8)
            text = open(os.path.join(wd, "e_results.dat")).read()
        np.savez_compressed(os.path.join(OUT, name + ".npz"), variant=variant, rows=rows, text=text, ref_seconds=a[1], **extra,
                            keys=np.array(sorted(f"{k}={v}" for k, v in kv.items())))
        print(f"{name}: {len(rows)} estimate calls, {int(rows[:, 5].sum())} eigenrays, reference {a[1]:.1f} s")
        print(rows)


if __name__ == "__main__":
    np.set_printoptions(linewidth=200, precision=9, suppress=True)
    main(sys.argv[1:] or list(CASES))
