set -x
N=$1
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r1q_config2_gpus$N.json 2> gpurun_out/r1q_config2_gpus$N.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --workload config5 --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/r1q_config5_gpus$N.json 2> gpurun_out/r1q_config5_gpus$N.err
tail -c 300 gpurun_out/r1q_config5_gpus$N.err
for w in config2 config5; do python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r1q_${w}_gpus$N.json").read().strip().splitlines()[-1])
    print("$w gpus=$N", "rays/s %.1f steps/s %.4g ms/pass %.1f %s e2e %.1f occ %.3f rays/gpu %d" % (d["value"], d["rk4_steps_per_sec"], d["ms_per_step"], d["scaling"], d["e2e"]["value"], d["config"]["lane_occupancy"], d["config"]["rays_per_gpu"]))
except Exception as e: print("$w failed", e)
PY
done
