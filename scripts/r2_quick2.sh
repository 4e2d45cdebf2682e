#!/bin/bash
# whole GPU suite (the scan is everywhere) + short bench
TAG=${1:-r2q}; OUT=gpurun_out/$TAG; mkdir -p $OUT
echo "== pytest"; timeout 1500 python -m pytest tests -m gpu -q -x --deselect tests/test_gpu_fullsize.py 2>&1 | tail -5 | tee $OUT/pytest.txt
echo "== bench"; env $2 timeout 600 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-secondary --e2e-steps 3 2> $OUT/bench.err | tee $OUT/bench.json | python -c "
import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['phase_ms'], 'frac', d['roofline']['frac'], 'e2e', d['e2e']['ms_per_step'], 'renum', d['renumbered_samples']['ms_per_step'])"
tail -3 $OUT/bench.err
