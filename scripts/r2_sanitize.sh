#!/bin/bash
# compute-sanitizer over scripts/sanitize_all.py (every kernel family at small sizes); logs -> gpurun_out/san/
OUT=gpurun_out/san; mkdir -p $OUT
for tool in memcheck racecheck initcheck; do
  echo "== $tool"
  timeout 900 compute-sanitizer --tool $tool --error-exitcode 7 python scripts/sanitize_all.py > $OUT/$tool.log 2>&1
  echo "exit $?" >> $OUT/$tool.log
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY|sanitize run complete|^exit" $OUT/$tool.log | tail -4
done
# the block-per-column conversion (shared-memory network with warp-level synchronisation of the short strides)
echo "== racecheck, block sort"
MPB200_SLAB_BLOCK_SORT=1 timeout 900 compute-sanitizer --tool racecheck --error-exitcode 7 python scripts/sanitize_all.py > $OUT/racecheck_block_sort.log 2>&1
echo "exit $?" >> $OUT/racecheck_block_sort.log
grep -E "RACECHECK SUMMARY|sanitize run complete|^exit" $OUT/racecheck_block_sort.log | tail -3
