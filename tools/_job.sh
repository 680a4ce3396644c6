mkdir -p gpurun_out/r2y
timeout 1200 python -m pytest tests -x -q -m gpu > gpurun_out/r2y/pytest_all.log 2>&1
tail -3 gpurun_out/r2y/pytest_all.log
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r2y/smoke.log 2>&1; tail -1 gpurun_out/r2y/smoke.log
( time python bench.py > gpurun_out/r2y/bench_default.json 2> gpurun_out/r2y/bench_default.err ) 2> gpurun_out/r2y/bench_default.time
( time python bench.py --impl reference > gpurun_out/r2y/bench_reference.json 2> gpurun_out/r2y/bench_reference.err ) 2> gpurun_out/r2y/bench_reference.time
cat gpurun_out/r2y/bench_default.time gpurun_out/r2y/bench_reference.time | grep real
python - <<'PY'
import json
for f in ("bench_default","bench_reference"):
    try:
        d=json.loads([l for l in open(f"gpurun_out/r2y/{f}.json") if l.startswith("{")][-1])
        print(f, d.get("value"), d.get("e2e",{}).get("value"), d.get("ms_per_step"), d.get("roofline",{}).get("frac"), d.get("cpu_baseline"), d.get("parity_check"), d.get("clocks"))
    except Exception as e: print(f,"FAILED",e)
PY
