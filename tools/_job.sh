mkdir -p gpurun_out/r2w
run() { # lib tag workload kernel extra
  TR_LIB=$1 timeout 600 python bench.py --workload $3 --steps 2 --warmup 1 --moves-per-point 1000 --no-cpu-baseline --kernel $4 --parity-chains 2 $5 $6 2> gpurun_out/r2w/$3_$2.err | grep '^{' > gpurun_out/r2w/$3_$2.json
  python - gpurun_out/r2w/$3_$2.json $3 $2 <<'PY'
import json, sys
try:
    d = json.load(open(sys.argv[1])); p = d.get("parity_check", {})
    print(sys.argv[2], sys.argv[3], d["roofline"]["kernel"], d["config"]["chains_per_gpu"], round(d["value"] / 1e6, 1), "M moves/s e2e", round(d["e2e"]["value"] / 1e6, 1), "frac", round(d["roofline"]["frac"], 4), "refwork", round(d["roofline"]["reference_work_equivalent"]["frac"], 4), "parity", p.get("decisions_equal"), p.get("max_rel_err"))
except Exception as e:
    print(sys.argv[2], sys.argv[3], "FAILED", e)
PY
}
run exp_libs/lib_pad.so pad h2o20 0
run transrot_b200/libtransrot_b200.so base h2o20 0
run exp_libs/lib_pad.so pad sample8 0
run transrot_b200/libtransrot_b200.so base sample8 0
run exp_libs/lib_pad.so pad2 h2o20 0
