# round 2, GPU run 11: the quarter rule (cost > quarter_alpha % of the average lane work) on one GPU's share of the 8- / 4- / 2-way split of config 5
set -x
P=gpurun_out/r2l
B="python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-e2e --workload config5"
run() { name=$1; shard=$2; shift 2; env "$@" timeout 400 $B --shard-of $shard > ${P}_$name.json 2> ${P}_$name.err; python - <<PY
import json
try:
    d=json.loads(open("${P}_$name.json").read().strip().splitlines()[-1]); print("RESULT $name", round(d["ms_per_step"]), "ms", round(d["rk4_steps_per_sec"]/1e6), "Msteps/s occ", d["config"]["lane_occupancy"], d["config"].get("schedule"))
except Exception as e: print("RESULT $name failed", e)
PY
}
run s8_default 0/8 A=1
run s8_qa150 0/8 GEOAC_B200_QUARTER_ALPHA=150
run s4_default 0/4 A=1
run s4_qoff 0/4 GEOAC_B200_QUARTER=0
run s2_default 0/2 A=1
