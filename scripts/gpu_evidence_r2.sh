# end-of-round-2 evidence run (1 GPU): tests, bench lines, reference arm, near-threshold listing, ncu launch list + section captures
set -x
P=gpurun_out/r2z
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active --format=csv > ${P}_smi.csv
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
( time timeout 1500 python -m pytest tests -m gpu -q -s ) > ${P}_pytest_gpu.log 2>&1; tail -5 ${P}_pytest_gpu.log
timeout 900 python bench.py > ${P}_bench_config2.json 2> ${P}_bench_config2.err; tail -c 900 ${P}_bench_config2.json
timeout 900 python bench.py --impl reference > ${P}_bench_reference_arm.json 2> ${P}_bench_reference_arm.err; tail -c 500 ${P}_bench_reference_arm.json
timeout 600 python bench.py --workload config1 --no-strong > ${P}_bench_config1.json 2>/dev/null; tail -c 300 ${P}_bench_config1.json
timeout 900 python bench.py --workload config3 --steps 2 --warmup 1 --no-strong > ${P}_bench_config3.json 2>/dev/null; tail -c 300 ${P}_bench_config3.json
timeout 900 python bench.py --workload config4 --steps 2 --warmup 1 --no-cpu-baseline --no-strong > ${P}_bench_config4.json 2>/dev/null; tail -c 300 ${P}_bench_config4.json
timeout 600 python bench.py --workload config5 --steps 1 --warmup 0 --no-cpu-baseline --no-e2e --shard-of 0/8 > ${P}_bench_config5_rank0of8.json 2>/dev/null; tail -c 400 ${P}_bench_config5_rank0of8.json
timeout 300 python tools/list_near_threshold.py --workload config2 --every 100 --limit 60 > ${P}_near_threshold_config2.txt 2>&1; head -5 ${P}_near_threshold_config2.txt
B="python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-strong"
timeout 300 $B > ${P}_plain.log 2>&1 && timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file ${P}_launches_config2.csv $B > ${P}_ncu_launch.log 2>&1
SEC="--section SpeedOfLight --section ComputeWorkloadAnalysis --section MemoryWorkloadAnalysis --section SchedulerStats --section WarpStateStats --section InstructionStats --section LaunchStats --section Occupancy"
cap() { name=$1; shift; timeout 600 ncu $SEC --clock-control none -k regex:trace_kernel -s 1 -c 1 "$@" > ${P}_ncu_$name.txt 2>&1; tail -3 ${P}_ncu_$name.txt; }
cap Eq3D_config2_slice python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-strong --no-e2e --rays-cap 86400
cap EqGlobal_config3_slice python bench.py --workload config3 --steps 1 --warmup 1 --no-cpu-baseline --no-strong --no-e2e --rays-cap 57000
cap Eq3DRD_config4_full_occupancy python bench.py --workload config4 --steps 1 --warmup 1 --no-cpu-baseline --no-strong --no-e2e --rays-cap 40000
cap EqGlobalRD_config5_full_occupancy python bench.py --workload config5 --steps 1 --warmup 1 --no-cpu-baseline --no-strong --no-e2e --rays-cap 40000 --ray-limit 200
ls -la gpurun_out | grep r2z
