run() { # name lib args...
  name=$1; lib=$2; shift 2
  TR_LIB=$lib timeout 200 python bench.py --steps 2 --warmup 1 --no-cpu-baseline "$@" 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$name', d['config']['chains_per_gpu'], round(d['value']/1e6,2), round(d['roofline']['frac'],4))"
}
for w in sample8 nh4cl32 mixed64; do
run base_$w transrot_b200/libtransrot_b200.so --workload $w --moves-per-point 500
run ns1_$w exp_libs/lib_ns1.so --workload $w --moves-per-point 500
done
run base_h2o256 transrot_b200/libtransrot_b200.so --workload h2o256 --moves-per-point 100
run ns1_h2o256 exp_libs/lib_ns1.so --workload h2o256 --moves-per-point 100
run base_h2o20_t16 transrot_b200/libtransrot_b200.so --workload h2o20 --moves-per-point 500 --threads-per-chain 16
run ns1_h2o20_t16 exp_libs/lib_ns1.so --workload h2o20 --moves-per-point 500 --threads-per-chain 16
