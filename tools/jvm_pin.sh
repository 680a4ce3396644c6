#!/bin/bash
# jvm_pin.sh — pin the oracle (and, on a GPU box, the CUDA path) to OUTPUT OF THE REAL REFERENCE.
#
# Needs a JDK (javac + java) and the reference checkout ($REF, default /root/reference).  Neither exists in the build
# image nor on the GPU boxes of this pool (bench.py records the probe in "jdk_probe"), so this recipe has never run:
# it is the ready-to-run procedure for the first machine that has both.  Nothing is copied from the reference: its
# sources are compiled where they lie, outputs go to oracle/_ref/ (git-ignored).
#
#   tools/jvm_pin.sh                 # compile, run every case of tools/jvm_compare.py, compare, write oracle/_ref/pin_report.md
#   REF=/path/to/TransRot tools/jvm_pin.sh
set -euo pipefail
ROOT=$(cd "$(dirname "$0")/.." && pwd)
REF=${REF:-/root/reference}
OUT=${OUT:-$ROOT/oracle/_ref}
JAR=$REF/src/resources/commons-math3-3.6.1.jar
command -v javac >/dev/null || { echo "jvm_pin: no javac on PATH (parity stays unpinned)"; exit 3; }
command -v java  >/dev/null || { echo "jvm_pin: no java on PATH (parity stays unpinned)"; exit 3; }
[ -f "$JAR" ] || { echo "jvm_pin: $JAR not found (set REF=...)"; exit 3; }
mkdir -p "$OUT/classes" "$OUT/cases"
java -version 2>&1 | head -1 > "$OUT/jdk_version.txt"
# the reference has no build script (Transrot.iml is an IntelliJ module): Main-Class main.Main, one jar on the class path
javac -cp "$JAR" -d "$OUT/classes" "$REF"/src/main/*.java
python "$ROOT/tools/jvm_compare.py" make-cases "$OUT/cases" --dbase "$REF/src/dbase.txt" --input "$REF/src/Input.xyz"
for c in "$OUT"/cases/*/; do
  seed=$(cat "$c/seed.txt")
  args=(-c "$c/config.txt" -d "$c/dbase.txt" -o "$c/out" -s "$seed")
  [ -f "$c/Input.xyz" ] && args+=(-i "$c/Input.xyz")
  rm -rf "$c/out"; mkdir -p "$c/out"
  ( cd "$c" && java -cp "$OUT/classes:$JAR" main.Main "${args[@]}" > "$c/stdout.txt" 2> "$c/stderr.txt" ) || echo "case $c: JVM exit $?"
done
python "$ROOT/tools/jvm_compare.py" compare "$OUT/cases" --report "$OUT/pin_report.md"
