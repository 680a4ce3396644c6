run() { # name env -- args
  name=$1; shift
  envs=(); while [ "$1" != "--" ]; do envs+=("$1"); shift; done; shift
  env "${envs[@]}" timeout 200 python bench.py --steps 2 --warmup 1 --no-cpu-baseline "$@" 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$name', d['config']['chains_per_gpu'], round(d['value']/1e6,2), 'M moves/s frac', round(d['roofline']['frac'],4))"
}
run nh4cl32_team X=1 -- --workload nh4cl32 --moves-per-point 500
run nh4cl32_crew TR_FORCE_CREW=1 -- --workload nh4cl32 --moves-per-point 500
run nh4cl32_crew_bps3 TR_FORCE_CREW=1 TR_CREW_BPS=3 -- --workload nh4cl32 --moves-per-point 500
run mixed64_crew TR_FORCE_CREW=1 -- --workload mixed64 --moves-per-point 500
run sample8_crew X=1 -- --workload sample8 --moves-per-point 500
