set -x
timeout 900 python -m pytest tests/test_eigenray.py -m gpu -q -s 2>&1 | tail -45 > gpurun_out/r1z_pytest_eigenray.log; tail -4 gpurun_out/r1z_pytest_eigenray.log
