set -x
B="python bench.py --workload config5s --steps 1 --warmup 1 --no-cpu-baseline --rays-cap 4096 --bounces 0"
timeout 300 $B > gpurun_out/r1k_plain.log 2>&1 && timeout 900 ncu --section SourceCounters --section LaunchStats --section MemoryWorkloadAnalysis --clock-control none --import-source on -k regex:trace_kernel -s 1 -c 1 -o gpurun_out/r1k_prof_grd $B > gpurun_out/r1k_ncu.log 2>&1
tail -2 gpurun_out/r1k_ncu.log; python -c "
import json
d=json.loads(open('gpurun_out/r1k_plain.log').read().strip().splitlines()[-1]); print(d['ms_per_step'], d['rk4_steps_per_sec'], d['config']['lane_occupancy'], d['config']['rk4_steps_per_pass_per_gpu'])"
