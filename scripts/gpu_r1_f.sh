set -x
B="python bench.py --workload config4s --steps 1 --warmup 1 --no-cpu-baseline"
timeout 300 $B > gpurun_out/r1f_plain.log 2>&1 && timeout 1500 ncu --set full --clock-control none --import-source on -k regex:trace_kernel -s 1 -c 1 -o gpurun_out/r1f_prof_rd $B > gpurun_out/r1f_ncu_full.log 2>&1
tail -3 gpurun_out/r1f_ncu_full.log
