#!/bin/bash
# bench_configs.sh [out_dir] [extra bench.py flags...] - every BASELINE configuration at FULL schedule length (20 000 moves per point)
# on one GPU, with the oracle parity check of the chains just timed.  One JSON line per configuration into <out_dir>/cfg_<name>.json.
out=${1:-gpurun_out}; shift
mkdir -p "$out"
for w in sample8 h2o20 nh4cl32 mixed64 h2o256; do
  timeout 900 python bench.py --workload $w --steps 2 --warmup 1 --cpu-seconds 5 "$@" 2> "$out/cfg_$w.err" | grep '^{' > "$out/cfg_$w.json"
  python - "$out/cfg_$w.json" $w <<'PY'
import json, sys
try:
    d = json.load(open(sys.argv[1]))
    p = d.get("parity_check", {})
    print(sys.argv[2], d["roofline"]["kernel"], d["config"]["chains_per_gpu"], "chains", round(d["value"] / 1e6, 1), "M moves/s, e2e", round(d["e2e"]["value"] / 1e6, 1),
          "frac", round(d["roofline"]["frac"], 4), "ref-work frac", round(d["roofline"]["reference_work_equivalent"]["frac"], 4), "ms", round(d["ms_per_step"], 1),
          "parity", p.get("decisions_equal"), p.get("max_rel_err"))
except Exception as e:
    print(sys.argv[2], "FAILED", e)
PY
done
