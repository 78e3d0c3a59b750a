# round 2, GPU run 7: the whole GPU test suite on the current build; config 2 / 3 with the cost scout tracing every 1st / 4th / 8th ray
set -x
P=gpurun_out/r2g
( time timeout 1200 python -m pytest tests -m gpu -q -x ) > ${P}_pytest.log 2>&1; tail -6 ${P}_pytest.log
for S in 1 4 8; do
 for W in config2 config3; do
  GEOAC_B200_SCOUT_STRIDE=$S timeout 300 python bench.py --workload $W --steps 3 --warmup 2 --no-cpu-baseline --no-strong --no-e2e > ${P}_${W}_s$S.json 2> ${P}_${W}_s$S.err
  python - <<PY
import json
d=json.loads(open("${P}_${W}_s$S.json").read().strip().splitlines()[-1]); print("RESULT $W scout_stride $S", round(d["ms_per_step"],1), "ms", round(d["rk4_steps_per_sec"]/1e9,3), "Gsteps/s frac", round(d["roofline"]["frac"],4), "occ", d["config"]["lane_occupancy"])
PY
 done
done
