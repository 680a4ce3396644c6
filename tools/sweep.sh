run() { # name lib args...
  name=$1; lib=$2; shift 2
  TR_LIB=$lib timeout 120 python bench.py --steps 3 --warmup 1 --moves-per-point 1000 --no-cpu-baseline "$@" 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$name', d['config']['chains_per_gpu'], round(d['value']/1e9,4), round(d['ms_per_step'],2))"
}
run base transrot_b200/libtransrot_b200.so
run kouter exp_libs/lib_kout.so
