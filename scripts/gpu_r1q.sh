set -x
timeout 900 python bench.py > gpurun_out/r1z_bench_config2.json 2> gpurun_out/r1z_bench_config2.err; tail -c 400 gpurun_out/r1z_bench_config2.json
timeout 900 python bench.py --workload config1 > gpurun_out/r1z_bench_config1.json 2>/dev/null; tail -c 200 gpurun_out/r1z_bench_config1.json
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
