mkdir -p gpurun_out/r2t
timeout 1200 python -m pytest tests -x -q -m gpu > gpurun_out/r2t/pytest_all.log 2>&1
tail -3 gpurun_out/r2t/pytest_all.log
bash tools/bench_configs.sh gpurun_out/r2t/cfg
