"""GPU tests at BASELINE.json's full sizes, through size-independent properties (no oracle: the CPU reference would
need hours): determinism, independence of the claim order / scheduling mode, record sanity, sharded == unsharded, and
the error paths of the C ABI.  Parity proper (against the reference's golden vectors) is tests/test_gpu_parity.py."""
import ctypes as C
import os
import sys

import numpy as np
import pytest

import geoac_b200 as g
from geoac_b200 import abi, api, sharding
from tests import util

sys.path.insert(0, util.ROOT)
import bench            # noqa: E402  (workload definitions only)

pytestmark = pytest.mark.gpu


def _checksum(out):
    """Order-sensitive checksum of checksums over every record field (bit patterns, so NaN-safe and exact)."""
    M = (1 << 64) - 1
    h = 1469598103934665603
    for k in ("status", "n_steps"):
        h = ((h * 1099511628211) & M) ^ (int(out[k].astype(np.int64).sum()) & M)
    bits = np.ascontiguousarray(out["rec"]).view(np.uint64)
    w = (np.arange(bits.size, dtype=np.uint64).reshape(bits.shape) | np.uint64(1))
    with np.errstate(over="ignore"):
        x = int(np.bitwise_xor.reduce((bits * w).ravel()))
    return ((h * 1099511628211) & M) ^ x


def test_config2_full_grid_properties(monkeypatch):
    """216 000 rays (config 2): two runs are bitwise identical, the longest-ray-first schedule does not change a bit, every
    slot is accounted for, arrivals carry sane records, and the azimuthal structure of the stratified problem holds."""
    _, th_deg, ph_deg, th, ph = bench.workload_angles("config2")
    tr, p = bench.setup_tracer("config2", 0)
    a = tr.trace(th, ph)
    occ = tr.last_lane_occupancy()
    b = tr.trace(th, ph)
    assert _checksum(a) == _checksum(b)
    tr2, _ = bench.setup_tracer("config2", 0)
    tr2.set_knob("lpt", 0)
    c = tr2.trace(th, ph)
    assert tr2.last_kernel_launches() == 1 and _checksum(a) == _checksum(c)
    assert occ > 0.93                                                   # the scheduling pass keeps the warps full
    st = a["status"]
    assert set(np.unique(st)) <= {abi.ST_NONE, abi.ST_ARRIVAL, abi.ST_BREAK, abi.ST_LIMIT}
    assert (st[:, 0] != abi.ST_NONE).all()                              # every ray produced a first-segment outcome
    arr = st == abi.ST_ARRIVAL
    assert 400000 < arr.sum() <= st.size
    tt = a["rec"][abi.F_TRAVELTIME][arr]
    assert np.isfinite(a["rec"][:, arr]).all() and (tt > 0).all() and (a["n_steps"][arr] > 2).all()
    # a slot after a BREAK / LIMIT is never filled
    ended = (st[:, :-1] != abi.ST_ARRIVAL)
    assert (st[:, 1:][ended] == abi.ST_NONE).all()
    # celerity of every arrival (range / travel time) is physical: 0.2 ... 0.36 km/s
    rng = np.hypot(a["rec"][0][arr], a["rec"][1][arr])
    cel = rng / tt
    assert cel.min() > 0.15 and cel.max() < 0.40


def test_rngdep_scale_run_is_schedule_independent(monkeypatch):
    """Range-dependent variant on a 50x50x300 grid, 10 000 rays: packet scheduling + longest-first order vs natural order,
    and two host-thread-driven contexts vs one: bitwise identical."""
    _, _, _, th, ph = bench.workload_angles("config4s")
    th, ph = th[::3].copy(), ph[::3].copy()
    tr, p = bench.setup_tracer("config4s", 0)
    tr.set_knob("lpt", 2)
    a = tr.trace(th, ph)
    tr.set_knob("lpt", 0)
    b = tr.trace(th, ph)
    assert _checksum(a) == _checksum(b)
    assert (a["status"][:, 0] != abi.ST_NONE).all() and (a["status"] == abi.ST_ARRIVAL).sum() > 1000
    tr2, _ = bench.setup_tracer("config4s", 0)
    c = sharding.trace_multi([tr, tr2], th, ph, block=256)
    assert _checksum(a) == _checksum(c)


def test_abi_error_paths():
    L = api.lib()
    st = C.c_int(0)
    assert not L.geoac_create(99, 0, C.byref(st)) and st.value == abi.GEOAC_ERR_BAD_ARG
    assert not L.geoac_create(abi.GEOAC_3D, 4096, C.byref(st)) and st.value == abi.GEOAC_ERR_BAD_ARG
    tr = g.Tracer(abi.GEOAC_3D, 0)
    one = np.zeros(1)
    rec = np.zeros((abi.NFIELDS, 1, 3)); s = np.zeros((1, 3), dtype=np.int32)
    dp, ip = C.POINTER(C.c_double), C.POINTER(C.c_int32)
    call = lambda n: L.geoac_trace(tr._h, n, one.ctypes.data_as(dp), one.ctypes.data_as(dp), rec.ctypes.data_as(dp), s.ctypes.data_as(ip), s.ctypes.data_as(ip))
    assert call(1) == abi.GEOAC_ERR_NO_ATMO and b"atmosphere" in L.geoac_last_error(tr._h)
    assert call(-1) == abi.GEOAC_ERR_BAD_ARG
    z, T, u, v, rho = g.load_met_1d(util.TOY)
    with pytest.raises(g.GeoAcError):
        tr.set_atmosphere_1d(z[::-1], T, u, v, rho)                    # altitudes must increase
    with pytest.raises(g.GeoAcError):
        tr.set_atmosphere_1d(z[:2], T[:2], u[:2], v[:2], rho[:2])      # too few levels
    with pytest.raises(g.GeoAcError):
        tr.set_atmosphere_3d(z, z, z, np.zeros((3, 3, 3)), np.zeros((3, 3, 3)), np.zeros((3, 3, 3)), np.zeros((3, 3, 3)))   # wrong variant
    rd = g.Tracer(abi.GEOAC_3D_RNGDEP, 0)
    with pytest.raises(g.GeoAcError):
        rd.set_atmosphere_1d(z, T, u, v, rho)
    p = tr.params
    p.bounces = -1
    with pytest.raises(g.GeoAcError):
        tr.params = p
    tr.set_atmosphere_1d(z, T, u, v, rho)
    assert call(0) == abi.GEOAC_OK and call(1) == abi.GEOAC_OK         # recovers after errors


def test_table_too_large_for_shared_memory_still_traces():
    """A profile with more levels than fit beside the lane records in shared memory falls back to reading the table through
    L1/L2 (TABLE_IN_SMEM = false) and gives the same records as the shared-memory path on the same spline."""
    z, T, u, v, rho = g.load_met_1d(util.TOY)
    th, ph = util.angles_rad(np.linspace(4, 40, 19), np.full(19, 20.0))
    tr = g.Tracer(abi.GEOAC_3D, 0)
    tr.set_atmosphere_1d(z, T, u, v, rho)
    a = tr.trace(th, ph)
    # same profile resampled on a 4x finer grid by the spline's own knots would change the spline; instead append levels
    # above the propagation ceiling, which the rays never see (vert_limit stays at the original top)
    zx = np.concatenate([z, z[-1] + 0.1 * np.arange(1, 2001)])
    ext = lambda f: np.concatenate([f, np.full(2000, f[-1])])
    big = g.Tracer(abi.GEOAC_3D, 0)
    big.set_atmosphere_1d(zx, ext(T), ext(u), ext(v), ext(rho))
    pb = big.params
    pb.vert_limit = tr.params.vert_limit
    big.params = pb
    b = big.trace(th, ph)
    assert np.array_equal(a["status"], b["status"]) and (np.abs(a["n_steps"] - b["n_steps"]) <= 1).all()
    m = a["status"] == abi.ST_ARRIVAL
    # the natural spline's end condition moves to the new top, which perturbs the slopes of the last original levels a little
    assert np.allclose(a["rec"][abi.F_TRAVELTIME][m], b["rec"][abi.F_TRAVELTIME][m], rtol=1e-6)
    # raypath / caustic capture works for such a profile too (the reference's WriteRays mode has no size limit)
    for t in (tr, big):
        q = t.params
        q.accum_per_segment = 1
        t.params = q
    pa, pb = tr.trace_paths(th, ph, 25, 2000, caustic_cap=16), big.trace_paths(th, ph, 25, 2000, caustic_cap=16)
    assert (np.abs(pa["path_rows"] - pb["path_rows"]) <= 1).all() and pa["path_rows"].min() > 50
    n0 = int(min(pa["path_rows"][0], pb["path_rows"][0])) - 2
    assert np.allclose(pa["path"][0, :n0, :3], pb["path"][0, :n0, :3], rtol=1e-5, atol=1e-6)
    assert np.array_equal(pa["caustic_rows"], pb["caustic_rows"])


@pytest.mark.parametrize("glob", [False, True])
def test_device_built_node_tables_match_set_slopes_multi(glob, monkeypatch):
    """geoac_set_atmosphere_3d builds the node tables on the device (one thread per column and quantity).  They must be
    bit for bit what Set_Slopes_Multi yields (G2S_MultiDimSpline3D.cpp:306-425 / G2S_GlobalMultiDimSpline3D.cpp:313-431,
    restated by the oracle, incl. the Global file's dfdt[i]-dfdt[i+1] slip) and what the host builder of the same library
    yields (knob host_tables = 1), on a non-uniform vertical axis so that every spacing-dependent term differs."""
    from oracle import pyoracle as po          # checker only
    from geoac_b200 import synth
    n0, n1, nz = 9, 11, 57
    ax0, ax1, axz, T, u, v, rho = (synth.config5_grid(n0, n1, nz) if glob else synth.config4_grid(n0, n1, nz))
    rng = np.random.default_rng(7)
    axz = np.cumsum(0.3 + rng.random(nz))                    # irregular levels
    T = T * (1.0 + 0.01 * rng.standard_normal(T.shape)); u = u + 1e-3 * rng.standard_normal(u.shape)
    variant = abi.GEOAC_GLOBAL_RNGDEP if glob else abi.GEOAC_3D_RNGDEP
    tr = g.Tracer(variant, 0)
    tr.set_atmosphere_3d(ax0, ax1, axz, T, u, v, rho)
    tuv, rh = tr.grid_tables(n0, n1, nz)
    ref = po.atmo3d_slopes(po.atmo3d(glob, ax0, ax1, axz, T, u, v, rho), (n0, n1, nz))
    for F in range(3):
        for w in range(4):
            assert np.array_equal(tuv[..., 6 * F + w].view(np.uint64), ref[F, w].view(np.uint64)), (F, w)
    assert np.array_equal(rh[..., 0].view(np.uint64), ref[3, 0].view(np.uint64))
    assert np.array_equal(rh[..., 1].view(np.uint64), ref[3, 1].view(np.uint64))
    th = g.Tracer(variant, 0)
    th.set_knob("host_tables", 1)
    th.set_atmosphere_3d(ax0, ax1, axz, T, u, v, rho)
    tuv_h, rh_h = th.grid_tables(n0, n1, nz)
    assert np.array_equal(tuv.view(np.uint64), tuv_h.view(np.uint64)) and np.array_equal(rh.view(np.uint64), rh_h.view(np.uint64))


@pytest.mark.parametrize("workload", ["config1", "config2", "config3"])
def test_absorption_polynomials_match_the_full_model(workload, monkeypatch):
    """The stratified kernels take the Sutherland-Bass coefficient from per-interval polynomials of its smooth factors
    (core.cuh: sb_alpha_1d; the cancellation staircase of the classical term stays literal).  Against the full model
    evaluated at every step (knob sbpoly = 0) the accumulated absorption must agree to 1e-11 and every other output must
    be bit-identical -- on ToyAtmo (no interval flagged) and on the config-3 profile (a few intervals fall back to the full
    model because the piecewise-linear temperature kinks inside them)."""
    _, _, _, th, ph = bench.workload_angles(workload)
    step = max(1, len(th) // 3000)
    th, ph = th[::step].copy(), ph[::step].copy()
    tr, p = bench.setup_tracer(workload, 0)
    tr.set_knob("sbpoly", 1)
    a = tr.trace(th, ph)
    tr.set_knob("sbpoly", 0)
    b = tr.trace(th, ph)
    assert np.array_equal(a["status"], b["status"]) and np.array_equal(a["n_steps"], b["n_steps"])
    m = a["status"] == abi.ST_ARRIVAL
    assert m.sum() > 100
    for f in range(abi.NFIELDS):
        if f == abi.F_ATTEN:
            rel = np.abs(a["rec"][f][m] - b["rec"][f][m]) / np.abs(b["rec"][f][m])
            assert rel.max() < 1e-11, rel.max()
        else:
            assert np.array_equal(a["rec"][f][m].view(np.uint64), b["rec"][f][m].view(np.uint64)), f
    # a changed frequency rebuilds the table
    q = tr.params
    q.freq = 2.5
    tr.params = q
    tr.set_knob("sbpoly", 1)
    c = tr.trace(th[:200], ph[:200])
    tr.set_knob("sbpoly", 0)
    d = tr.trace(th[:200], ph[:200])
    mm = c["status"] == abi.ST_ARRIVAL
    rel = np.abs(c["rec"][abi.F_ATTEN][mm] - d["rec"][abi.F_ATTEN][mm]) / np.abs(d["rec"][abi.F_ATTEN][mm])
    assert rel.max() < 1e-11 and (c["rec"][abi.F_ATTEN][mm] > 5 * a["rec"][abi.F_ATTEN][:200][mm]).all()
