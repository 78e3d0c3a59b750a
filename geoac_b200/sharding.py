"""Multi-GPU sharding of a launch-angle batch: rays are independent, so the flattened ray list
(phi-major, theta-minor -- the mains' loop order) is split across devices in interleaved blocks, the atmosphere is
replicated on every device, and the per-device records are put back in the original order on the host.
There is NO collective on the data path (north_star: "no NCCL is needed on the path").

Two drivers use the same partition:
  * one process per GPU (bench.py under torchrun; tests/test_sharding.py with gloo ranks on CPU);
  * one process driving several contexts from host threads (`trace_multi`), which is what a C++ front end would do with
    one geoac_ctx per device (include/geoac_b200.h: "calls on one ctx are serialised by the caller").
"""
import threading

import numpy as np

from . import abi

SHARD_BLOCK = 4096      # rays per block: lifetimes vary smoothly with theta, so interleaved blocks balance the load


def shard_indices(n_rays, rank, world, block=SHARD_BLOCK):
    """Indices (ascending) of the rays rank `rank` of `world` traces."""
    if not (0 <= rank < world):
        raise ValueError("rank out of range")
    idx = np.arange(n_rays, dtype=np.int64)
    return idx[(idx // block) % world == rank]


def empty_records(n_rays, n_rec):
    return {"rec": np.zeros((abi.NFIELDS, n_rays, n_rec)), "status": np.zeros((n_rays, n_rec), dtype=np.int32),
            "n_steps": np.zeros((n_rays, n_rec), dtype=np.int32)}


def merge_shards(n_rays, n_rec, shards):
    """shards: iterable of (indices, records) -> records of the whole batch in the original ray order.
    Raises if the shards do not cover every ray exactly once."""
    out = empty_records(n_rays, n_rec)
    seen = np.zeros(n_rays, dtype=np.int32)
    for idx, rec in shards:
        seen[idx] += 1
        out["rec"][:, idx, :] = rec["rec"]
        out["status"][idx] = rec["status"]
        out["n_steps"][idx] = rec["n_steps"]
    if not (seen == 1).all():
        raise ValueError("shards must cover every ray exactly once")
    return out


def trace_multi(tracers, theta, phi, block=SHARD_BLOCK):
    """Trace one batch on several contexts (one per device, same variant / atmosphere / parameters) and return the merged
    records -- a thin binding over geoac_trace_multi (include/geoac_b200.h): the partition, the host threads, the pinned staging
    and the merge by ray index all live in the library, which is what a C++ front end calls too.  `block` other than the
    library's GEOAC_SHARD_BLOCK is honoured with the host-side partition below (tests exercise ragged blocks with it).
    The result is bitwise identical to tracing the batch on one device."""
    from . import api
    if block == SHARD_BLOCK:
        return api.trace_multi(tracers, theta, phi)
    theta = np.ascontiguousarray(theta, dtype=np.float64)
    phi = np.ascontiguousarray(phi, dtype=np.float64)
    world = len(tracers)
    n_rec = tracers[0].params.bounces + 1
    parts = [None] * world
    errors = []

    def work(r):
        try:
            idx = shard_indices(len(theta), r, world, block)
            parts[r] = (idx, tracers[r].trace(theta[idx], phi[idx]))
        except Exception as e:          # surfaced to the caller below
            errors.append(e)

    threads = [threading.Thread(target=work, args=(r,)) for r in range(world)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    if errors:
        raise errors[0]
    return merge_shards(len(theta), n_rec, parts)
