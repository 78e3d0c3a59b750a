set -x
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
timeout 1200 python -m pytest tests -m gpu -q -s 2>&1 | tail -60 > gpurun_out/pytest_gpu.log; tail -40 gpurun_out/pytest_gpu.log
