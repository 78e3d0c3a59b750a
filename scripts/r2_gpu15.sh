# round 2, final 1-GPU check of the shipped build: smoke, the whole GPU suite, the default bench line (config 2 + config-5 strong sub-record)
set -x
P=gpurun_out/r2z
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
( time timeout 1500 python -m pytest tests -m gpu -q -s ) > ${P}_pytest_gpu.log 2>&1; tail -5 ${P}_pytest_gpu.log
timeout 900 python bench.py > ${P}_bench_config2.json 2> ${P}_bench_config2.err; tail -c 600 ${P}_bench_config2.json
