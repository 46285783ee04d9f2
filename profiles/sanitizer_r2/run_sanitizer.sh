#!/bin/bash
# compute-sanitizer pass over every kernel family.  Run on a GPU box:
#   gpurun --timeout 2400 -- 'bash profiles/sanitizer_r2/run_sanitizer.sh'
# One python process per tool runs all cases of benchmarks/sanitizer_cases.py; logs land in gpurun_out/sanitizer_r2/
# (copied to profiles/sanitizer_r2/ afterwards).
set -u
OUT=gpurun_out/sanitizer_r2
mkdir -p $OUT
SAN=/usr/local/cuda/bin/compute-sanitizer
for tool in ${TOOLS:-racecheck memcheck synccheck}; do
  extra=""
  [ $tool = racecheck ] && extra="--racecheck-report all"
  t0=$(date +%s)
  timeout ${SAN_TIMEOUT:-1200} $SAN --tool $tool $extra --print-limit 30 python benchmarks/sanitizer_cases.py ${CASES:-} > $OUT/${tool}.log 2>&1
  rc=$?
  echo "$tool rc=$rc $(( $(date +%s) - t0 )) s cases: $(grep -c '^case .* done' $OUT/${tool}.log) | $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' $OUT/${tool}.log | tail -1)"
done | tee $OUT/summary.txt
