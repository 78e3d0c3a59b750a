set -x
timeout 1200 python -m pytest tests/test_eigenray.py -m gpu -q -x -s 2>&1 | tail -40 > gpurun_out/r1l_pytest_eig.log; tail -40 gpurun_out/r1l_pytest_eig.log
