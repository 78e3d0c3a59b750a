set -x
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -5 
for b in 512 384 256; do
GEOAC_B200_BLOCK=$b timeout 600 python bench.py --no-cpu-baseline > gpurun_out/r1c_bench_$b.json 2> gpurun_out/r1c_bench_$b.err; python - <<PY
import json
d=json.loads(open("gpurun_out/r1c_bench_$b.json").read().strip().splitlines()[-1])
print($b, d["value"], d["rk4_steps_per_sec"], d["roofline"]["frac"], d["config"]["lane_occupancy"], d["e2e"]["value"])
PY
done
B="python bench.py --steps 1 --warmup 1 --no-cpu-baseline"
timeout 300 $B > gpurun_out/r1c_plain.log 2>&1 && timeout 1200 ncu --set full --clock-control none --import-source on -k regex:trace_kernel -s 1 -c 1 -o gpurun_out/r1c_prof $B > gpurun_out/r1c_ncu_full.log 2>&1
tail -3 gpurun_out/r1c_ncu_full.log
