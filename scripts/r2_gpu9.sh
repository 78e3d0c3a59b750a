# round 2, GPU run 9: quarter claims for the very longest long-region packets (config 5, rank 0 of 8 on one GPU)
set -x
P=gpurun_out/r2j
timeout 600 python -m pytest tests -m gpu -q -x -k "neutral or rngdep_scale or config5_full or globalrngdep" > ${P}_pytest.log 2>&1; tail -3 ${P}_pytest.log
B="python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-e2e --workload config5 --shard-of 0/8"
run() { name=$1; shift; env "$@" timeout 400 $B > ${P}_$name.json 2> ${P}_$name.err; python - <<PY
import json
try:
    d=json.loads(open("${P}_$name.json").read().strip().splitlines()[-1]); print("RESULT $name", round(d["ms_per_step"]), "ms", round(d["rk4_steps_per_sec"]/1e6), "Msteps/s occ", d["config"]["lane_occupancy"], d["config"].get("schedule"))
except Exception as e: print("RESULT $name failed", e)
PY
}
run q_default A=1
run q_off GEOAC_B200_QUARTER=0
run q_pct65 GEOAC_B200_LONG_SM_PCT=65
run q_pct80 GEOAC_B200_LONG_SM_PCT=80
run q_pct65_alpha70 GEOAC_B200_LONG_SM_PCT=65 GEOAC_B200_LONG_ALPHA=70
