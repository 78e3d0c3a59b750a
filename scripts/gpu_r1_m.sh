set -x
for cfg in "libgeoac_b200.so 384" "libgeoac_b200.so 448" "libgeoac_b200.so 320" "libgeoac_b200_exp.so 384" "libgeoac_b200_exp.so 448"; do set -- $cfg
GEOAC_B200_LIB=$1 GEOAC_B200_BLOCK=$2 timeout 600 python bench.py --no-cpu-baseline > gpurun_out/r1m_$1_$2.json 2> gpurun_out/r1m_$1_$2.err; python - <<PY
import json
d=json.loads(open("gpurun_out/r1m_$1_$2.json").read().strip().splitlines()[-1])
print("$1 block=$2", d["value"], d["rk4_steps_per_sec"], d["roofline"]["frac"], d["config"]["lane_occupancy"])
PY
done
