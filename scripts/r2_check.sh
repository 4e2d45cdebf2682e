#!/bin/bash
# round-2 one-GPU check: new tests first, then the whole GPU suite, smoke, and the full bench line
TAG=${1:-r2b}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > $OUT/gpu.txt 2>&1
echo "== new tests"; timeout 1200 python -m pytest tests/test_gpu_fullsize.py -m gpu -x -q 2>&1 | tail -15 | tee $OUT/pytest_fullsize.txt
echo "== pytest -m gpu"; timeout 1500 python -m pytest tests -m gpu -q --deselect tests/test_gpu_fullsize.py 2>&1 | tail -15 | tee $OUT/pytest_gpu.txt
echo "== smoke"; timeout 300 python __graft_entry__.py smoke 2>&1 | tail -5 | tee $OUT/smoke.txt
echo "== bench"; timeout 1200 python bench.py --steps 30 --warmup 5 > $OUT/bench.json 2> $OUT/bench.err; tail -5 $OUT/bench.err; cat $OUT/bench.json
