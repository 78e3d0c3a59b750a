# round-1 evidence run (1 GPU): tests, bench lines of configs 1-4, reference arm, ncu launch list + full capture
set -x
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active --format=csv > gpurun_out/r1_final_smi.csv
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 1200 python -m pytest tests -m gpu -q -s 2>&1 | tail -120 > gpurun_out/r1_pytest_gpu_final.log; tail -3 gpurun_out/r1_pytest_gpu_final.log
timeout 900 python bench.py > gpurun_out/r1_bench_config2_final.json 2> gpurun_out/r1_bench_config2_final.err; tail -c 600 gpurun_out/r1_bench_config2_final.json
timeout 900 python bench.py --impl reference > gpurun_out/r1_bench_reference_arm_final.json 2> gpurun_out/r1_bench_reference_arm_final.err; tail -c 500 gpurun_out/r1_bench_reference_arm_final.json
timeout 900 python bench.py --workload config1 > gpurun_out/r1_bench_config1_final.json 2>/dev/null
timeout 900 python bench.py --workload config3 --steps 2 --warmup 1 > gpurun_out/r1_bench_config3_final.json 2>/dev/null; tail -c 300 gpurun_out/r1_bench_config3_final.json
timeout 900 python bench.py --workload config4 --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/r1_bench_config4_final.json 2>/dev/null; tail -c 300 gpurun_out/r1_bench_config4_final.json
B="python bench.py --steps 2 --warmup 1 --no-cpu-baseline"
timeout 300 $B > gpurun_out/r1_final_plain.log 2>&1 && timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r1_final_launches.csv $B > gpurun_out/r1_final_ncu_launch.log 2>&1
B2="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --rays-cap 86400"
timeout 300 $B2 > gpurun_out/r1_final_plain2.log 2>&1 && timeout 1200 ncu --set full --clock-control none --import-source on -k regex:trace_kernel -s 1 -c 1 -o gpurun_out/r1_final_prof $B2 > gpurun_out/r1_final_ncu_full.log 2>&1
tail -2 gpurun_out/r1_final_ncu_full.log
