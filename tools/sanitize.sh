#!/bin/bash
# compute-sanitizer over the hot kernels: memcheck, racecheck, synccheck on tools/sanitize_cases.py.
# Usage (on a GPU box):  tools/sanitize.sh [out_dir] [case ...]      Logs: <out_dir>/sanitize_<tool>.log + a summary.
out=${1:-gpurun_out}; shift
cases="$@"
mkdir -p "$out"
CS=/usr/local/cuda/bin/compute-sanitizer
: > "$out/sanitize_summary.txt"
for tool in memcheck racecheck synccheck; do
  extra=""
  [ $tool = racecheck ] && extra="--racecheck-report all"
  timeout ${SAN_TIMEOUT:-600} $CS --tool $tool $extra --print-limit 20 --log-file "$out/sanitize_$tool.log" python tools/sanitize_cases.py $cases > "$out/sanitize_$tool.out" 2>&1
  rc=$?
  echo "== $tool: exit $rc; $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' "$out/sanitize_$tool.log" | tail -1)" >> "$out/sanitize_summary.txt"
  cat "$out/sanitize_$tool.out" >> "$out/sanitize_summary.txt"
done
cat "$out/sanitize_summary.txt"
