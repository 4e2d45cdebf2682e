#!/bin/bash
TAG=${1:-r2c3}; OUT=gpurun_out/$TAG; mkdir -p $OUT
echo "== k3 tests"; timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py tests/test_gpu_edge_cases.py -m gpu -q -x -k "k3 or high_dim or c3 or all_pairs or single_sample" 2>&1 | tail -8 | tee $OUT/pytest.txt
echo "== C3"; timeout 600 python bench_configs.py --configs C3 2>&1 | tail -2 | cut -c1-1200 | tee $OUT/c3.json
echo "== C3 one-sided"; MPB200_TC_FULL=1 timeout 600 python bench_configs.py --configs C3 2>&1 | tail -1 | cut -c1-500
