#!/bin/bash
for v in "X=1" "MPB200_FLOOR_I32=1" "MPB200_FLOOR_I32=1 MPB200_FLOOR_BANDS=8"; do
  env $v python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-secondary --e2e-steps 1 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$v floor ms', d['roofline']['write_pattern_floor_ms'], 'fill', d['roofline']['kernel_ms'])"
done
