# round 2, GPU run 10: share of the SMs given to the long-region launch x step multiple of the cost scout (config 5, rank 0 of 8)
set -x
P=gpurun_out/r2k
B="python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-e2e --workload config5 --shard-of 0/8"
run() { name=$1; shift; env "$@" timeout 400 $B > ${P}_$name.json 2> ${P}_$name.err; python - <<PY
import json
try:
    d=json.loads(open("${P}_$name.json").read().strip().splitlines()[-1]); print("RESULT $name", round(d["ms_per_step"]), "ms", round(d["rk4_steps_per_sec"]/1e6), "Msteps/s occ", d["config"]["lane_occupancy"], d["config"].get("schedule"))
except Exception as e: print("RESULT $name failed", e)
PY
}
run pct90 GEOAC_B200_LONG_SM_PCT=90
run pct80_sc32 GEOAC_B200_LONG_SM_PCT=80 GEOAC_B200_SCOUT_COARSE=32
run pct90_sc32 GEOAC_B200_LONG_SM_PCT=90 GEOAC_B200_SCOUT_COARSE=32
run pct80_sc64 GEOAC_B200_LONG_SM_PCT=80 GEOAC_B200_SCOUT_COARSE=64
run pct80_sc32_alpha150 GEOAC_B200_LONG_SM_PCT=80 GEOAC_B200_SCOUT_COARSE=32 GEOAC_B200_LONG_ALPHA=150
