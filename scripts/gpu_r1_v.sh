set -x
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | grep -v "max rel diff\|listed" | tail -4
run() { name=$1; shift; timeout 900 python bench.py --workload $name --no-cpu-baseline "$@" > gpurun_out/r1v_$name.json 2> gpurun_out/r1v_$name.err; tail -c 400 gpurun_out/r1v_$name.err; python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r1v_$name.json").read().strip().splitlines()[-1])
    print("$name", "rays/s %.1f steps/s %.4g ms/pass %.1f frac %.3f (survey fig %.3f) occ %.3f e2e %.1f" % (d["value"], d["rk4_steps_per_sec"], d["ms_per_step"], d["roofline"]["frac"], d["roofline"]["achieved_survey_figure"]/d["roofline"]["peak"], d["config"]["lane_occupancy"], d["e2e"]["value"]))
except Exception as e: print("$name failed", e)
PY
}
run config4 --steps 1 --warmup 1
run config5 --steps 1 --warmup 0 --rays-cap 100000
