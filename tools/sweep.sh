run() { # name env...
  name=$1; shift
  env "$@" timeout 120 python bench.py --steps 3 --warmup 1 --moves-per-point 1000 --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$name', d['config']['chains_per_gpu'], round(d['value']/1e9,4), round(d['ms_per_step'],2))"
}
run base X=1
run carve50 TR_CREW_CARVEOUT=50
run carve62 TR_CREW_CARVEOUT=62
run carve75 TR_CREW_CARVEOUT=75
run carve100 TR_CREW_CARVEOUT=100
