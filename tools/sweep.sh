run() { # name lib args...
  name=$1; lib=$2; shift 2
  TR_LIB=$lib timeout 120 python bench.py --steps 3 --warmup 1 --moves-per-point 1000 --no-cpu-baseline "$@" 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$name', d['config']['chains_per_gpu'], round(d['value']/1e9,4), round(d['ms_per_step'],2))"
}
run minb2_2072 transrot_b200/libtransrot_b200.so --chains-per-gpu 2072
run minb1_2072 exp_libs/lib_m1.so --chains-per-gpu 2072
run minb2_4096 transrot_b200/libtransrot_b200.so
run minb1_4096 exp_libs/lib_m1.so
