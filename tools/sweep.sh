for bps in 2 3 4 6 8; do
  TR_CREW_BPS=$bps timeout 120 python bench.py --steps 2 --warmup 1 --moves-per-point 500 --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('bps=$bps', d['value']/1e9, d['roofline']['frac'])"
done
for c in 1 2 4 8 16; do
  TR_CREW_C=$c timeout 120 python bench.py --steps 2 --warmup 1 --moves-per-point 500 --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('C=$c', d['value']/1e9, d['roofline']['frac'])"
done
ncu --set full --clock-control none --import-source on -k regex:crew_kernel -c 1 -f -o gpurun_out/prof_crew1 python bench.py --steps 1 --warmup 0 --moves-per-point 100 --no-cpu-baseline > gpurun_out/crew1_ncu.log 2>&1
