set -x
timeout 900 python -m pytest tests/test_gpu_scale.py tests/test_gpu_parity.py -m gpu -q -x 2>&1 | tail -3
run() { env "$@" timeout 600 python bench.py --workload config2 --no-cpu-baseline > gpurun_out/r1q_tmp.json 2> gpurun_out/r1q_tmp.err; python -c "
import json,sys; d=json.loads(open('gpurun_out/r1q_tmp.json').read().strip().splitlines()[-1]); print('RESULT', sys.argv[1:], d['value'], d['rk4_steps_per_sec'], d['ms_per_step'], d['roofline']['frac'], d['config']['lane_occupancy'])" "$@"; }
run GEOAC_B200_COSTSHIFT=3
run GEOAC_B200_COSTSHIFT=2
run GEOAC_B200_COSTSHIFT=4
run GEOAC_B200_COSTSHIFT=0
run GEOAC_B200_COSTSHIFT=3 GEOAC_B200_PACKET=0
cp gpurun_out/r1q_tmp.err gpurun_out/r1q_last.err
