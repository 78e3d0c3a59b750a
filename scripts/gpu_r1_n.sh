set -x
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | grep -v "max rel diff\|listed" | tail -3
run() { name=$1; shift; timeout 900 python bench.py --workload $name --no-cpu-baseline "$@" > gpurun_out/r1n_$name.json 2> gpurun_out/r1n_$name.err; tail -c 400 gpurun_out/r1n_$name.err; python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r1n_$name.json").read().strip().splitlines()[-1])
    print("$name", "rays/s %.1f steps/s %.4g ms/pass %.1f frac %.3f (survey fig %.3f) occ %.3f e2e %.1f" % (d["value"], d["rk4_steps_per_sec"], d["ms_per_step"], d["roofline"]["frac"], d["roofline"]["achieved_survey_figure"]/d["roofline"]["peak"], d["config"]["lane_occupancy"], d["e2e"]["value"]))
except Exception as e: print("$name failed", e)
PY
}
run config2 --steps 3 --warmup 3
run config3 --steps 1 --warmup 1
run config4 --steps 1 --warmup 1
B="python bench.py --steps 1 --warmup 1 --no-cpu-baseline"
timeout 300 $B > gpurun_out/r1n_plain.log 2>&1 && timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r1n_launches.csv $B > gpurun_out/r1n_ncu_launch.log 2>&1
grep -E "scout|trace_kernel" gpurun_out/r1n_launches.csv | cut -d, -f5,15 | cut -c1-160 | head -8
