# round 2, GPU run 4: cooperative kernel with the TMA cell cache -- correctness subset, then config 5 (rank 0 of 8, and in full) and config 4
set -x
P=gpurun_out/r2d
timeout 900 python -m pytest tests -m gpu -q -x -k "neutral or schedule or rngdep or config4_full or config5_full or golden or multi or library" > ${P}_pytest.log 2>&1; tail -4 ${P}_pytest.log
B="python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-e2e"
run() { name=$1; wl=$2; shard=$3; shift 3; env "$@" timeout 400 $B --workload $wl $shard > ${P}_$name.json 2> ${P}_$name.err; python - <<PY
import json
try:
    d=json.loads(open("${P}_$name.json").read().strip().splitlines()[-1]); print("RESULT $name", round(d["ms_per_step"]), "ms", round(d["rk4_steps_per_sec"]/1e6), "Msteps/s frac", round(d["roofline"]["frac"],3), "occ", d["config"]["lane_occupancy"], d["config"].get("schedule"))
except Exception as e: print("RESULT $name failed", e)
PY
}
run c5s8_default config5 "--shard-of 0/8" A=1
run c5s8_allcoop config5 "--shard-of 0/8" GEOAC_B200_LONG_ALPHA=1 GEOAC_B200_LONG_SM_PCT=90
run c5s8_pct90 config5 "--shard-of 0/8" GEOAC_B200_LONG_SM_PCT=90
run c5s8_serial_excl config5 "--shard-of 0/8" GEOAC_B200_LONG_WIDTH=32
run c5full_allcoop config5 "" GEOAC_B200_LONG_ALPHA=1 GEOAC_B200_LONG_SM_PCT=90 GEOAC_B200_LPT=2
run c4_allcoop config4 "" GEOAC_B200_LONG_ALPHA=1 GEOAC_B200_LONG_SM_PCT=90 GEOAC_B200_LPT=2
run c4_default config4 "" A=1
