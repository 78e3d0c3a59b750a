# round 2, GPU run 13: every rank's share of the 8-way config-5 split, one after another on one GPU (block 4096, then 1024)
set -x
P=gpurun_out/r2n
B="python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-e2e --workload config5"
for BLK in 4096 1024; do
for R in 0 1 2 3 4 5 6 7; do
  GEOAC_BENCH_SHARD_BLOCK=$BLK timeout 300 $B --shard-of $R/8 > ${P}_b${BLK}_r$R.json 2> ${P}_b${BLK}_r$R.err
  python - <<PY
import json
try:
    d=json.loads(open("${P}_b${BLK}_r$R.json").read().strip().splitlines()[-1]); print("RESULT block $BLK rank $R", round(d["ms_per_step"]), "ms steps", d["config"]["rk4_steps_per_pass_per_gpu"], d["config"].get("schedule"))
except Exception as e: print("RESULT block $BLK rank $R failed", e)
PY
done
done
