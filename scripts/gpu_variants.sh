#!/bin/bash
# compare kernel variants selected by environment variables: quick benches, no CPU baseline
OUT=gpurun_out/${1:-variants}; mkdir -p $OUT
for v in "MPB200_FILL_U=1" "MPB200_FILL_U=2" "MPB200_FILL_U=4" "MPB200_FILL_U=2 MPB200_NO_CLASSIFY=1"; do
  echo "== $v"; env $v timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --e2e-steps 1 2>&1 | tail -1 | python -c "
import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['phase_ms'])"
done
