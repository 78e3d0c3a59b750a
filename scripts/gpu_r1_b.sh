set -x
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -15 > gpurun_out/r1b_pytest.log; tail -8 gpurun_out/r1b_pytest.log
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/r1b_bench.json 2> gpurun_out/r1b_bench.err; tail -c 2500 gpurun_out/r1b_bench.json; tail -5 gpurun_out/r1b_bench.err
B="python bench.py --workload midgrid --steps 2 --warmup 1 --no-cpu-baseline"
timeout 300 $B > gpurun_out/r1b_plain.log 2>&1 && timeout 1200 ncu --set full --clock-control none --import-source on -k regex:trace_kernel -s 1 -c 1 -o gpurun_out/r1b_prof $B > gpurun_out/r1b_ncu_full.log 2>&1
tail -3 gpurun_out/r1b_ncu_full.log
