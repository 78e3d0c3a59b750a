set -x
B2="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --rays-cap 86400"
timeout 300 $B2 > gpurun_out/r1r_plain2.log 2>&1 && timeout 1200 ncu --set full --clock-control none --import-source on -k regex:trace_kernel -s 1 -c 1 -o gpurun_out/r1r_prof $B2 > gpurun_out/r1r_ncu_full.log 2>&1
tail -2 gpurun_out/r1r_ncu_full.log
