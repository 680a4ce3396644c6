# moves/s of every BASELINE configuration (short schedules; the headline run is bench.py's default)
for w in sample8 h2o20 nh4cl32 mixed64 h2o256; do
  mpp=500; [ $w = h2o256 ] && mpp=100
  timeout 300 python bench.py --workload $w --steps 2 --warmup 1 --moves-per-point $mpp --no-cpu-baseline 2>gpurun_out/cfg_$w.err | tee gpurun_out/cfg_$w.json | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$w', d['config']['chains_per_gpu'], 'chains', round(d['value']/1e6,2), 'M moves/s  frac', round(d['roofline']['frac'],4), 'ms', round(d['ms_per_step'],1), 'e2e', round(d['e2e']['value']/1e6,2))"
done
