set -x
timeout 900 python -m pytest tests/test_gpu_scale.py tests/test_gpu_parity.py -m gpu -q -x 2>&1 | tail -3
for B in 384 512; do
GEOAC_B200_BLOCK=$B timeout 600 python bench.py --workload config2 --no-cpu-baseline > gpurun_out/r1p_config2_$B.json 2> gpurun_out/r1p_config2_$B.err; python -c "
import json,sys; d=json.loads(open('gpurun_out/r1p_config2_$B.json').read().strip().splitlines()[-1]); print($B, d['value'], d['rk4_steps_per_sec'], d['ms_per_step'], d['roofline']['frac'])"
done
