# round 2, GPU run 16: ordering cost dilated over the inclination neighbours -- the slow rank (7) and a normal one (0) of the 8-way config-5 split; rank 3 of 4
set -x
P=gpurun_out/r2r
timeout 600 python -m pytest tests -m gpu -q -x -k "neutral or rngdep_scale" > ${P}_pytest.log 2>&1; tail -3 ${P}_pytest.log
B="python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-e2e --workload config5"
for S in 7/8 0/8 3/4; do
  timeout 300 $B --shard-of $S > ${P}_tmp.json 2> ${P}_tmp.err
  python - <<PY
import json
try:
    d=json.loads(open("${P}_tmp.json").read().strip().splitlines()[-1]); print("RESULT dilate rank $S", round(d["ms_per_step"]), "ms", d["config"].get("schedule"))
except Exception as e: print("RESULT rank $S failed", e)
PY
done
