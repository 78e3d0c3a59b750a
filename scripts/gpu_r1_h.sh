set -x
B="python bench.py --workload config4s --steps 1 --warmup 1 --no-cpu-baseline --rays-cap 4096 --bounces 0"
timeout 300 $B > gpurun_out/r1h_plain.log 2>&1 && timeout 900 ncu --section SourceCounters --section WarpStateStats --section LaunchStats --section Occupancy --section MemoryWorkloadAnalysis --section ComputeWorkloadAnalysis --section SchedulerStats --clock-control none --import-source on -k regex:trace_kernel -s 1 -c 1 -o gpurun_out/r1h_prof_rd $B > gpurun_out/r1h_ncu.log 2>&1
tail -3 gpurun_out/r1h_ncu.log; python -c "
import json
d=json.loads(open('gpurun_out/r1h_plain.log').read().strip().splitlines()[-1]); print(d['ms_per_step'], d['rk4_steps_per_sec'], d['config']['lane_occupancy'])"
