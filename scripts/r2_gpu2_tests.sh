# round 2, GPU run 2: the GPU test suite with the scale-parity tests, then config 5 in full once (per-ray step counts)
set -x
P=gpurun_out/r2b
timeout 1000 python -m pytest tests -m gpu -q -s -x > ${P}_pytest_gpu.log 2>&1; tail -5 ${P}_pytest_gpu.log
sed -i 's/r2a_config5_steps/r2b_config5_steps/' scripts/r2_config5_steps.py
timeout 420 python scripts/r2_config5_steps.py > ${P}_config5_steps.log 2>&1; tail -20 ${P}_config5_steps.log
