set -x
timeout 900 python -m pytest tests/test_gpu_scale.py tests/test_gpu_parity.py -m gpu -q -x 2>&1 | tail -8 > gpurun_out/r1n_pytest.log; tail -8 gpurun_out/r1n_pytest.log
timeout 600 python bench.py --workload config2 --no-cpu-baseline > gpurun_out/r1n_config2.json 2> gpurun_out/r1n_config2.err; tail -c 1800 gpurun_out/r1n_config2.json
GEOAC_B200_SBPOLY=0 timeout 600 python bench.py --workload config2 --no-cpu-baseline > gpurun_out/r1n_config2_nopoly.json 2> gpurun_out/r1n_config2_nopoly.err; tail -c 600 gpurun_out/r1n_config2_nopoly.json
timeout 600 python bench.py --workload config3 --no-cpu-baseline --steps 2 --warmup 1 > gpurun_out/r1n_config3.json 2> gpurun_out/r1n_config3.err; tail -c 900 gpurun_out/r1n_config3.json
