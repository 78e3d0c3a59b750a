# end-of-round-1 evidence run (1 GPU): tests, bench lines, reference arm, eigenray / ingest timings, ncu launch list + full capture
set -x
P=gpurun_out/r1z
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active --format=csv > ${P}_smi.csv
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 1500 python -m pytest tests -m gpu -q -s 2>&1 | tail -150 > ${P}_pytest_gpu.log; tail -3 ${P}_pytest_gpu.log
timeout 900 python bench.py > ${P}_bench_config2.json 2> ${P}_bench_config2.err; tail -c 700 ${P}_bench_config2.json
timeout 900 python bench.py --impl reference > ${P}_bench_reference_arm.json 2> ${P}_bench_reference_arm.err; tail -c 500 ${P}_bench_reference_arm.json
timeout 900 python bench.py --workload config1 > ${P}_bench_config1.json 2>/dev/null; tail -c 300 ${P}_bench_config1.json
timeout 900 python bench.py --workload config3 --steps 2 --warmup 1 > ${P}_bench_config3.json 2>/dev/null; tail -c 300 ${P}_bench_config3.json
timeout 900 python bench.py --workload config4 --steps 2 --warmup 1 --no-cpu-baseline > ${P}_bench_config4.json 2>/dev/null; tail -c 300 ${P}_bench_config4.json
timeout 600 python scripts/bench_eigenray.py 64 4 > ${P}_eigenray.json 2> ${P}_eigenray.err; cat ${P}_eigenray.json
B="python bench.py --steps 2 --warmup 1 --no-cpu-baseline"
timeout 300 $B > ${P}_plain.log 2>&1 && timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file ${P}_launches.csv $B > ${P}_ncu_launch.log 2>&1
B2="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --rays-cap 86400"
timeout 300 $B2 > ${P}_plain2.log 2>&1 && timeout 1200 ncu --set full --clock-control none --import-source on -k regex:trace_kernel -s 1 -c 1 -o ${P}_prof $B2 > ${P}_ncu_full.log 2>&1
tail -2 ${P}_ncu_full.log
