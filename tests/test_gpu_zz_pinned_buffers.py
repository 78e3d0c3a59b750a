"""geoac_trace into page-locked buffers from geoac_host_alloc (what a C++ front end hands over instead of the reference's
`new double*[...]` arrays, Code/GeoAc/GeoAc.Interface.cpp:53-58) gives the records of a trace into ordinary memory, bit for bit."""
import numpy as np
import pytest

import geoac_b200 as g
from geoac_b200 import abi, api
from tests import util

pytestmark = pytest.mark.gpu


def test_trace_into_pinned_buffers_is_bitwise_identical():
    tr = g.Tracer(abi.GEOAC_3D, 0)
    tr.set_atmosphere_1d(*api.load_met_1d(util.TOY))
    n_rec = tr.params.bounces + 1
    th = np.deg2rad(np.linspace(2.0, 40.0, 96))
    ph = np.deg2rad(np.linspace(0.0, 300.0, 96))
    want = tr.trace(th, ph)
    bufs = {"rec": api.PinnedArray((abi.NFIELDS, len(th), n_rec), np.float64), "status": api.PinnedArray((len(th), n_rec), np.int32),
            "n_steps": api.PinnedArray((len(th), n_rec), np.int32)}
    got = tr.trace(th, ph, out={k: v.array for k, v in bufs.items()})
    for k in ("rec", "status", "n_steps"):
        assert np.array_equal(got[k], want[k], equal_nan=(k == "rec")), k
    for v in bufs.values():
        v.close()
    tr.close()
