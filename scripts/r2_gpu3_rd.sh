# round 2, GPU run 3: range-dependent scheduling (packet grouping, exclusive long-region launch) -- correctness subset, then
# config 5 as rank 0 of an 8-way split (what one GPU of the N = 8 run does) under several knob settings, config 5 in full, config 4
set -x
P=gpurun_out/r2c
timeout 600 python -m pytest tests -m gpu -q -x -k "neutral or schedule or rngdep or config4_full or config5_full or golden" > ${P}_pytest.log 2>&1; tail -4 ${P}_pytest.log
B="python bench.py --workload config5 --steps 1 --warmup 0 --no-cpu-baseline --no-e2e"
run() { name=$1; shift; env "$@" timeout 400 $B --shard-of 0/8 > ${P}_c5s8_$name.json 2> ${P}_c5s8_$name.err; python - <<PY
import json
try:
    d=json.loads(open("${P}_c5s8_$name.json").read().strip().splitlines()[-1]); print("RESULT $name", round(d["ms_per_step"]), "ms", round(d["rk4_steps_per_sec"]/1e6), "Msteps/s occ", d["config"]["lane_occupancy"], d["config"].get("schedule"))
except Exception as e: print("RESULT $name failed", e)
PY
}
run default A=1
run old GEOAC_B200_RD_GROUP=0 GEOAC_B200_EXCLUSIVE=0
run theta_noexcl GEOAC_B200_EXCLUSIVE=0
run quarter GEOAC_B200_LONG_WIDTH=8
run alpha50 GEOAC_B200_LONG_ALPHA=50
timeout 500 $B > ${P}_c5full.json 2> ${P}_c5full.err; tail -c 600 ${P}_c5full.json
timeout 300 python bench.py --workload config4 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > ${P}_c4.json 2> ${P}_c4.err; tail -c 600 ${P}_c4.json
