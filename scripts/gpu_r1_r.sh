for P in 0 1; do
GEOAC_B200_PACKET=$P timeout 600 python bench.py --no-cpu-baseline > gpurun_out/r1r_$P.json 2> gpurun_out/r1r_$P.err; python - <<PY
import json
d=json.loads(open("gpurun_out/r1r_$P.json").read().strip().splitlines()[-1])
print("packet=$P config2", d["value"], d["rk4_steps_per_sec"], d["roofline"]["frac"], d["config"]["lane_occupancy"])
PY
GEOAC_B200_PACKET=$P timeout 600 python bench.py --no-cpu-baseline --workload config3 --steps 1 --warmup 1 > gpurun_out/r1r_c3_$P.json 2> gpurun_out/r1r_c3_$P.err; python - <<PY
import json
d=json.loads(open("gpurun_out/r1r_c3_$P.json").read().strip().splitlines()[-1])
print("packet=$P config3", d["value"], d["rk4_steps_per_sec"], d["roofline"]["frac"], d["config"]["lane_occupancy"])
PY
done
