set -x
timeout 1200 python bench.py --workload config5 --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/r1u_config5_gpus1.json 2> gpurun_out/r1u_config5_gpus1.err; tail -c 300 gpurun_out/r1u_config5_gpus1.err
python - <<PY
import json
d=json.loads(open("gpurun_out/r1u_config5_gpus1.json").read().strip().splitlines()[-1])
print("config5 N=1", "rays/s %.1f steps/s %.4g ms/pass %.1f occ %.3f e2e %.1f steps %d arrivals %d" % (d["value"], d["rk4_steps_per_sec"], d["ms_per_step"], d["config"]["lane_occupancy"], d["e2e"]["value"], d["config"]["rk4_steps_per_pass_per_gpu"], d["config"]["arrival_records_per_pass"]))
PY
