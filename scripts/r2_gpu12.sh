# round 2, GPU run 12: how fast does the cooperative kernel (TMA cell cache) advance the very longest config-5 rays? (long region = packets above 2.5x the average lane work only)
set -x
P=gpurun_out/r2m
B="python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-e2e --workload config5 --shard-of 0/8"
env GEOAC_B200_LONG_WIDTH=8 GEOAC_B200_LONG_ALPHA=250 GEOAC_B200_LONG_SM_PCT=70 timeout 400 $B > ${P}_coop_top.json 2> ${P}_coop_top.err
python - <<PY
import json
d=json.loads(open("${P}_coop_top.json").read().strip().splitlines()[-1]); print("RESULT coop_top", round(d["ms_per_step"]), "ms", d["config"].get("schedule"))
PY
