#!/bin/bash
TAG=${1:-r2c4}; OUT=gpurun_out/$TAG; mkdir -p $OUT
echo "== lq tests"; timeout 900 python -m pytest tests/test_gpu_lq.py tests/test_gpu_edge_cases.py tests/test_gpu_fmt.py "tests/test_gpu_fullsize.py::test_full_size_c4_lq_tables_and_edges" -m gpu -q -x 2>&1 | tail -8 | tee $OUT/pytest.txt
echo "== C4"; timeout 600 python bench_configs.py --configs C4 2>&1 | tail -3 | tee $OUT/c4.json
if [ "$2" = "ncu" ]; then
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:lq_inball_kernel -s 1 -c 1 -o $OUT/prof_lq -f python bench_configs.py --configs C4 > $OUT/ncu.log 2>&1; tail -2 $OUT/ncu.log
fi
