"""The algorithmic flop counts bench.py's roofline uses are the ones the op-counting build of the per-ray code produces
(tests/flopcount): re-run the counter on one variant and compare with the committed table and with bench.py."""
import json
import os
import subprocess
import sys

from tests import util


def test_committed_counts_match_bench_and_a_fresh_run():
    rows = [json.loads(l) for l in open(os.path.join(util.ROOT, "tests", "flopcount", "flops.jsonl"))]
    committed = {r["variant"]: r["flops_per_step"] for r in rows}
    sys.path.insert(0, util.ROOT)
    import bench
    names = {bench.V2D: "2d", bench.V3D: "3d", bench.VGLOBAL: "global", bench.V3DRD: "3drngdep", bench.VGLOBALRD: "globalrngdep"}
    for v, n in names.items():
        assert abs(bench.ALGO_FLOPS_PER_STEP[v] - committed[n]) <= 0.01 * committed[n], (n, bench.ALGO_FLOPS_PER_STEP[v], committed[n])
    out = subprocess.run([sys.executable, os.path.join(util.ROOT, "tests", "flopcount", "run_flopcount.py")], capture_output=True, text=True, check=True)
    fresh = {r["variant"]: r["flops_per_step"] for r in map(json.loads, out.stdout.strip().splitlines())}
    for n, f in fresh.items():
        assert abs(f - committed[n]) <= 0.02 * committed[n], (n, f, committed[n])
        assert rows[0]["transcendental"] < 40            # libm calls count once each, not as their expansions
