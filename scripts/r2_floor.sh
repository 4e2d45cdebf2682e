#!/bin/bash
for b in 0 2 4 8 16 64; do
  MPB200_FLOOR_BANDS=$b python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-secondary --e2e-steps 1 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('bands $b floor ms', d['roofline']['write_pattern_floor_ms'], 'fill', d['roofline']['kernel_ms'])"
done
