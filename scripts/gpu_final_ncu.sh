set -x
P=gpurun_out/r1z
timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -3 > ${P}_pytest_gpu_final_tail.log; cat ${P}_pytest_gpu_final_tail.log
B="python bench.py --steps 2 --warmup 1 --no-cpu-baseline"
timeout 300 $B > ${P}_plain.log 2>&1 && timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file ${P}_launches.csv $B > ${P}_ncu_launch.log 2>&1
B2="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --rays-cap 86400"
timeout 300 $B2 > ${P}_plain2.log 2>&1 && timeout 1200 ncu --set full --clock-control none --import-source on -k regex:trace_kernel -s 1 -c 1 -o ${P}_prof $B2 > ${P}_ncu_full.log 2>&1
tail -2 ${P}_ncu_full.log
