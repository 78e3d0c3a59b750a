set -x
timeout 900 python -m pytest tests/test_gpu_scale.py tests/test_gpu_parity.py -m gpu -q -x 2>&1 | tail -3
run() { env "$@" timeout 600 python bench.py --workload ${WL:-config2} --no-cpu-baseline ${EXTRA} > gpurun_out/r1q_tmp.json 2> gpurun_out/r1q_tmp.err; python -c "
import json,sys; d=json.loads(open('gpurun_out/r1q_tmp.json').read().strip().splitlines()[-1]); print('RESULT', sys.argv[1:], d['value'], d['rk4_steps_per_sec'], d['ms_per_step'], d['roofline']['frac'], d['config']['lane_occupancy'])" "$@"; }
run A=1
WL=config3 EXTRA="--steps 2 --warmup 1" run A=1
