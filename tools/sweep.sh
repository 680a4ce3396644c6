run() { # name lib extra-env...
  name=$1; lib=$2; shift 2
  env "$@" TR_LIB=$lib timeout 120 python bench.py --steps 2 --warmup 1 --moves-per-point 500 --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$name', round(d['value']/1e9,4), round(d['roofline']['frac'],4))"
}
run base transrot_b200/libtransrot_b200.so X=1
run e4 exp_libs/lib_e4.so X=1
run e4_c8 exp_libs/lib_e4.so TR_CREW_C=8
run e7b3 exp_libs/lib_e7b3.so X=1
run e7b3_c10 exp_libs/lib_e7b3.so TR_CREW_C=10
run e14 exp_libs/lib_e14.so X=1
run base_c12 transrot_b200/libtransrot_b200.so TR_CREW_C=12
