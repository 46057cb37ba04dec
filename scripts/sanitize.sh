#!/bin/bash
# compute-sanitizer (memcheck / racecheck / synccheck) over one small run of every kernel family;
# logs go to the directory given as $1 (default gpurun_out/sanitizer).  The device-side spin waits
# are given 120 s (the sanitizer slows the persistent kernels down by 10-100x).
OUT=${1:-gpurun_out/sanitizer}
mkdir -p "$OUT"
export LJMD_SPIN_TIMEOUT_S=120
for tool in memcheck racecheck synccheck; do
  for c in cluster ordered tiles cells gr; do
    timeout 900 compute-sanitizer --tool $tool --print-limit 20 python scripts/sanitize_case.py $c \
        > "$OUT/${tool}_${c}.log" 2>&1
    echo "$tool $c rc=$? $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' "$OUT/${tool}_${c}.log" | tail -1) | $(grep -E '^(cluster|ordered|tiles|cells|gr):' "$OUT/${tool}_${c}.log" | tail -1)"
  done
done | tee "$OUT/summary.txt"
