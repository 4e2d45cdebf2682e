#!/bin/bash
# quick C2 iteration: table parity tests + a short bench (no secondary, no CPU baseline); extra env via $2
TAG=${1:-r2q}; OUT=gpurun_out/$TAG; mkdir -p $OUT
echo "== parity"; timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_edge_cases.py -m gpu -x -q -k "k1_k2 or k2_ or full_size or edges_and_points or pairs or radius or fused" 2>&1 | tail -5 | tee $OUT/pytest.txt
echo "== bench"; env $2 timeout 600 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-secondary --e2e-steps 3 2> $OUT/bench.err | tee $OUT/bench.json | python -c "
import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['phase_ms'], 'frac', d['roofline']['frac'], 'e2e', d['e2e']['ms_per_step'], 'renum', d['renumbered_samples']['ms_per_step'])"
tail -3 $OUT/bench.err
