set -x
N=$1
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r1z_config2_gpus$N.json 2> gpurun_out/r1z_config2_gpus$N.err
tail -c 300 gpurun_out/r1z_config2_gpus$N.err
python - <<PY
import json
d=json.loads(open("gpurun_out/r1z_config2_gpus$N.json").read().strip().splitlines()[-1])
print("config2 gpus=$N", d["value"], d["rk4_steps_per_sec"], d["ms_per_step"], d["scaling"], d["e2e"]["value"])
PY
