# round 2, GPU run 14: two-level cost scout -- the slow rank (7) and a normal one (0) of the 8-way config-5 split; neutrality tests
set -x
P=gpurun_out/r2o
timeout 600 python -m pytest tests -m gpu -q -x -k "neutral or rngdep_scale or golden" > ${P}_pytest.log 2>&1; tail -3 ${P}_pytest.log
B="python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-e2e --workload config5"
for R in 7 0 4; do
  timeout 300 $B --shard-of $R/8 > ${P}_r$R.json 2> ${P}_r$R.err
  python - <<PY
import json
try:
    d=json.loads(open("${P}_r$R.json").read().strip().splitlines()[-1]); print("RESULT refine rank $R", round(d["ms_per_step"]), "ms", d["config"].get("schedule"))
except Exception as e: print("RESULT rank $R failed", e)
PY
done
GEOAC_B200_REFINE=0 timeout 300 $B --shard-of 7/8 > ${P}_r7_norefine.json 2>/dev/null; python - <<PY
import json
d=json.loads(open("${P}_r7_norefine.json").read().strip().splitlines()[-1]); print("RESULT norefine rank 7", round(d["ms_per_step"]), "ms", d["config"].get("schedule"))
PY
