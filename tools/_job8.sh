N=$1
mkdir -p gpurun_out/r2u
for w in ${WL:-h2o20 mixed64 h2o256}; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --workload $w --steps 2 --warmup 1 --no-cpu-baseline --parity-chains 1 > gpurun_out/r2u/${w}_${N}gpu.out 2> gpurun_out/r2u/${w}_${N}gpu.err
  grep '^{' gpurun_out/r2u/${w}_${N}gpu.out > gpurun_out/r2u/${w}_${N}gpu.json
  python - gpurun_out/r2u/${w}_${N}gpu.json $w $N <<'PY'
import json, sys
try:
    d = json.load(open(sys.argv[1])); p = d.get("parity_check", {})
    print(sys.argv[2], sys.argv[3], "GPUs", d["roofline"]["kernel"], d["config"]["chains_per_gpu"], "chains/GPU", round(d["value"] / 1e6, 1), "M moves/s e2e", round(d["e2e"]["value"] / 1e6, 1), "ms", round(d["ms_per_step"], 1), "parity", p.get("decisions_equal"), d["config"].get("gather"))
except Exception as e:
    print(sys.argv[2], sys.argv[3], "FAILED", e)
PY
done
if [ "$N" = "2" ]; then timeout 600 python -m pytest tests -x -q -m gpu -k "two_gpu or multi_engine" > gpurun_out/r2u/pytest_2gpu.log 2>&1; tail -3 gpurun_out/r2u/pytest_2gpu.log; fi
