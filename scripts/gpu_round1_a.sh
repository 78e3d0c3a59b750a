set -x
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
timeout 900 python -m pytest tests -m gpu -q -s 2>&1 | tail -40 > gpurun_out/r1_pytest.log; tail -15 gpurun_out/r1_pytest.log
timeout 600 python bench.py > gpurun_out/r1_bench.json 2> gpurun_out/r1_bench.err; tail -c 3000 gpurun_out/r1_bench.json
timeout 600 python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/r1_bench_ref.json 2> gpurun_out/r1_bench_ref.err; tail -c 1500 gpurun_out/r1_bench_ref.json
B="python bench.py --workload midgrid --steps 2 --warmup 1 --no-cpu-baseline"
timeout 300 $B > gpurun_out/r1_plain.log 2>&1 && timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r1_launches.csv $B > gpurun_out/r1_ncu_launch.log 2>&1
timeout 300 $B > gpurun_out/r1_plain2.log 2>&1 && timeout 1200 ncu --set full --clock-control none --import-source on -k regex:trace_kernel -s 1 -c 1 -o gpurun_out/r1_prof $B > gpurun_out/r1_ncu_full.log 2>&1
tail -5 gpurun_out/r1_ncu_full.log
