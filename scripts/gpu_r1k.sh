set -x
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -15 > gpurun_out/r1k_pytest.log; tail -15 gpurun_out/r1k_pytest.log
timeout 600 python scripts/time_ingest.py > gpurun_out/r1k_ingest.json 2> gpurun_out/r1k_ingest.err; cat gpurun_out/r1k_ingest.json; tail -3 gpurun_out/r1k_ingest.err
