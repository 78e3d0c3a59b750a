set -x
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | grep -v "max rel diff\|listed" | tail -8
for cfg in "1 512" "0 512" "1 384" "1 256"; do set -- $cfg
GEOAC_B200_LPT=$1 GEOAC_B200_BLOCK=$2 timeout 600 python bench.py --no-cpu-baseline > gpurun_out/r1d_bench_$1_$2.json 2> gpurun_out/r1d_bench_$1_$2.err; python - <<PY
import json
d=json.loads(open("gpurun_out/r1d_bench_$1_$2.json").read().strip().splitlines()[-1])
print("lpt=$1 block=$2", d["value"], d["rk4_steps_per_sec"], d["roofline"]["frac"], d["config"]["lane_occupancy"], d["e2e"]["value"], d["gpu_launches"])
PY
done
