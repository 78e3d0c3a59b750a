"""Multi-GPU path on CPU (world_size 2, gloo): the ray-list partition and the host-side merge of geoac_b200/sharding.py,
driven exactly like bench.py drives them under torchrun -- one process per rank, no data-path collective, results
gathered on rank 0 -- with the CPU oracle standing in for the per-rank tracer (the product itself has no CPU path).
The merged records must be bitwise identical to an unsharded run ("results at 2/4/8 GPUs are bitwise
identical to 1 GPU")."""
import os
import socket

import numpy as np
import pytest

from geoac_b200 import abi, sharding
from tests import util


def test_partition_covers_every_ray_once():
    for n in (0, 1, 4095, 4096, 4097, 100000):
        for world in (1, 2, 3, 8):
            parts = [sharding.shard_indices(n, r, world) for r in range(world)]
            allidx = np.concatenate(parts) if parts else np.zeros(0, dtype=np.int64)
            assert len(allidx) == n and np.array_equal(np.sort(allidx), np.arange(n))
            if n >= world * sharding.SHARD_BLOCK:
                sizes = [len(p) for p in parts]
                assert max(sizes) - min(sizes) <= sharding.SHARD_BLOCK        # interleaved blocks keep the ranks balanced
    with pytest.raises(ValueError):
        sharding.shard_indices(10, 2, 2)


def test_merge_rejects_gaps_and_overlaps():
    rec = sharding.empty_records(2, 1)
    with pytest.raises(ValueError):
        sharding.merge_shards(4, 1, [(np.array([0, 1]), rec)])
    with pytest.raises(ValueError):
        sharding.merge_shards(2, 1, [(np.array([0, 1]), rec), (np.array([1, 0]), rec)])


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _rank_main(rank, world, port, theta, phi, block, ret):
    import torch.distributed as dist
    from oracle import pyoracle as po
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        at = po.atmo1d(False, *po.load_met_1d(util.TOY))
        p = po.default_params(abi.GEOAC_2D, at)
        p.bounces = 1
        idx = sharding.shard_indices(len(theta), rank, world, block)
        out = po.trace(abi.GEOAC_2D, at, p, theta[idx], phi[idx])
        dist.barrier()                                       # the only collective besides the result gather: timing fence
        gathered = [None] * world if rank == 0 else None
        dist.gather_object((idx, {k: out[k] for k in ("rec", "status", "n_steps")}), gathered, dst=0)
        if rank == 0:
            ret["merged"] = sharding.merge_shards(len(theta), p.bounces + 1, gathered)
    finally:
        dist.destroy_process_group()


def test_two_rank_gloo_run_matches_single_rank(oracle):
    import torch.multiprocessing as mp
    theta_deg = np.linspace(3.0, 40.0, 22)
    th, ph = util.angles_rad(theta_deg, np.full(len(theta_deg), -90.0))
    at = oracle.atmo1d(False, *oracle.load_met_1d(util.TOY))
    p = oracle.default_params(abi.GEOAC_2D, at)
    p.bounces = 1
    want = oracle.trace(abi.GEOAC_2D, at, p, th, ph)
    ctx = mp.get_context("spawn")
    ret = ctx.Manager().dict()
    port = _free_port()
    procs = [ctx.Process(target=_rank_main, args=(r, 2, port, th, ph, 4, ret)) for r in range(2)]     # block of 4 rays: 6 interleaved blocks
    for pr in procs:
        pr.start()
    for pr in procs:
        pr.join(300)
        assert pr.exitcode == 0
    got = ret["merged"]
    assert np.array_equal(got["status"], want["status"]) and np.array_equal(got["n_steps"], want["n_steps"])
    assert np.array_equal(got["rec"], want["rec"])
