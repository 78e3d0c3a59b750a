# round 2, GPU run 6: lanes per SM of the config-2 kernel (384 / 416 / 448 with explicit register budgets), config 5 in full with one
# CTA per SM (working set of the rays in flight inside L2), the tests touched since run 5
set -x
P=gpurun_out/r2f
timeout 900 python -m pytest tests -m gpu -q -x -k "config2_every or library or too_large or interactive or neutral or rngdep_scale" > ${P}_pytest.log 2>&1; tail -4 ${P}_pytest.log
for BLK in 384 416 448; do
  GEOAC_B200_BLOCK=$BLK timeout 300 python bench.py --workload config2 --steps 3 --warmup 2 --no-cpu-baseline --no-strong > ${P}_c2_$BLK.json 2> ${P}_c2_$BLK.err
  python - <<PY
import json
d=json.loads(open("${P}_c2_$BLK.json").read().strip().splitlines()[-1]); print("RESULT config2 block $BLK", round(d["ms_per_step"],1), "ms", round(d["rk4_steps_per_sec"]/1e9,3), "Gsteps/s frac", round(d["roofline"]["frac"],4), "occ", d["config"]["lane_occupancy"])
PY
done
B="python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-e2e"
run() { name=$1; wl=$2; shard=$3; shift 3; env "$@" timeout 500 $B --workload $wl $shard > ${P}_$name.json 2> ${P}_$name.err; python - <<PY
import json
try:
    d=json.loads(open("${P}_$name.json").read().strip().splitlines()[-1]); print("RESULT $name", round(d["ms_per_step"]), "ms", round(d["rk4_steps_per_sec"]/1e6), "Msteps/s frac", round(d["roofline"]["frac"],3), "occ", d["config"]["lane_occupancy"], d["config"].get("schedule"))
except Exception as e: print("RESULT $name failed", e)
PY
}
run c5full_1cta config5 "" GEOAC_B200_RD_CTAS=1
run c5s8_1cta config5 "--shard-of 0/8" GEOAC_B200_RD_CTAS=1
run c4_1cta config4 "" GEOAC_B200_RD_CTAS=1
