#!/bin/bash
for t in 4 8 16 32; do
  MPB200_FETCH_THREADS=$t python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-secondary --e2e-steps 6 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('threads $t e2e ms', d['e2e']['ms_per_step'])"
done
python scripts/e2e_breakdown.py 2>&1 | tail -12
