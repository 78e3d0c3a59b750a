for lib in libgeoac_b200.so libgeoac_b200_exp.so; do
GEOAC_B200_LIB=$lib timeout 600 python bench.py --no-cpu-baseline > gpurun_out/r1w_$lib.json 2> gpurun_out/r1w_$lib.err; python - <<PY
import json
d=json.loads(open("gpurun_out/r1w_$lib.json").read().strip().splitlines()[-1])
print("$lib config2", d["value"], d["rk4_steps_per_sec"], d["roofline"]["frac"], d["config"]["lane_occupancy"], d["ms_per_step"])
PY
done
GEOAC_B200_LIB=libgeoac_b200_exp.so timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x 2>&1 | tail -2
GEOAC_B200_LIB=libgeoac_b200_exp.so timeout 600 python bench.py --no-cpu-baseline --workload config3 --steps 1 --warmup 1 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('exp config3', d['value'], d['rk4_steps_per_sec'])"
timeout 600 python bench.py --no-cpu-baseline --workload config3 --steps 1 --warmup 1 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('base config3', d['value'], d['rk4_steps_per_sec'])"
