# A/B helper (under gpurun): run `bench.py --workload $WL` once per argument, each argument being a space-separated list of
# environment settings, and print one RESULT line per run.   e.g.  WL=config2 bash scripts/gpu_ab.sh "A=1" "GEOAC_B200_SBPOLY=0"
for envs in "$@"; do
  env $envs timeout 600 python bench.py --workload ${WL:-config2} --no-cpu-baseline ${EXTRA} > gpurun_out/ab_tmp.json 2> gpurun_out/ab_tmp.err
  python -c "
import json,sys; d=json.loads(open('gpurun_out/ab_tmp.json').read().strip().splitlines()[-1]); print('RESULT', sys.argv[1], round(d['value'],1), d['rk4_steps_per_sec'], round(d['ms_per_step'],2), d['roofline']['frac'], d['config']['lane_occupancy'])" "$envs"
done
