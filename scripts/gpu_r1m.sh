set -x
timeout 900 python scripts/bench_eigenray.py 64 4 > gpurun_out/r1m_eigenray.json 2> gpurun_out/r1m_eigenray.err; cat gpurun_out/r1m_eigenray.json; tail -5 gpurun_out/r1m_eigenray.err
