#!/bin/bash
# compute-sanitizer over the tiny-fixture workload; logs → gpurun_out/sanitizer_*.log
mkdir -p gpurun_out
CS=/usr/local/cuda/bin/compute-sanitizer
for tool in memcheck synccheck racecheck initcheck; do
  extra=""
  [ "$tool" = "racecheck" ] && extra="--racecheck-report all"
  [ "$tool" = "initcheck" ] && extra="--track-unused-memory no"
  ( timeout ${SAN_TIMEOUT:-900} $CS --tool $tool $extra --print-limit 20 python scripts/sanitizer_workload.py --small 2>&1 | tail -60 ) > gpurun_out/sanitizer_$tool.log
  echo "== $tool: $(grep -c 'SANITIZER_WORKLOAD_OK' gpurun_out/sanitizer_$tool.log) ok-marker, $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' gpurun_out/sanitizer_$tool.log | tail -1)"
done
