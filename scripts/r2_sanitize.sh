#!/bin/bash
# compute-sanitizer over scripts/sanitize_all.py (every kernel family at small sizes); logs -> gpurun_out/san/
OUT=gpurun_out/san; mkdir -p $OUT
for tool in memcheck racecheck initcheck; do
  echo "== $tool"
  timeout 900 compute-sanitizer --tool $tool --error-exitcode 7 python scripts/sanitize_all.py > $OUT/$tool.log 2>&1
  echo "exit $?" >> $OUT/$tool.log
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY|sanitize run complete|^exit" $OUT/$tool.log | tail -4
done
