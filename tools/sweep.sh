run() { # name args...
  name=$1; shift
  timeout 300 python bench.py --steps 2 --warmup 1 --no-cpu-baseline "$@" 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$name', d['config']['chains_per_gpu'], round(d['value']/1e6,2), 'M moves/s frac', round(d['roofline']['frac'],4))"
}
run nh4cl32_crew --workload nh4cl32 --moves-per-point 500
run nh4cl32_team32 --workload nh4cl32 --moves-per-point 500 --threads-per-chain 32
run mixed64_team32 --workload mixed64 --moves-per-point 500 --threads-per-chain 32
run sample8_team16 --workload sample8 --moves-per-point 500 --threads-per-chain 16
run h2o20_team16 --workload h2o20 --moves-per-point 500 --threads-per-chain 16
run h2o256_128 --workload h2o256 --moves-per-point 100 --threads-per-chain 128
run h2o256_256 --workload h2o256 --moves-per-point 100 --threads-per-chain 256
run h2o256_256_592 --workload h2o256 --moves-per-point 100 --threads-per-chain 256 --chains-per-gpu 592
run h2o256_128_888 --workload h2o256 --moves-per-point 100 --threads-per-chain 128 --chains-per-gpu 888
run h2o256_128_1184 --workload h2o256 --moves-per-point 100 --threads-per-chain 128 --chains-per-gpu 1184
