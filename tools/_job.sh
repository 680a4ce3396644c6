mkdir -p gpurun_out/r2p
timeout 1700 python -m pytest tests -x -q -m gpu > gpurun_out/r2p/pytest_all.log 2>&1
tail -5 gpurun_out/r2p/pytest_all.log
