run() { # name extra-env...
  name=$1; shift
  env "$@" timeout 120 python bench.py --steps 3 --warmup 1 --moves-per-point 1000 --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$name', round(d['value']/1e9,4), round(d['roofline']['frac'],4))"
}
run base X=1
run st2500_m2 TR_CREW_STAGGER=2500 TR_CREW_STAGGER_MODE=2
run st1500_m2 TR_CREW_STAGGER=1500 TR_CREW_STAGGER_MODE=2
run st3500_m2 TR_CREW_STAGGER=3500 TR_CREW_STAGGER_MODE=2
