#!/bin/bash
# One gpurun call: GPU tests, smoke, bench, ncu launch list + full capture of the top kernels.
# usage: scripts/gpu_check.sh [tag]
TAG=${1:-run}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > $OUT/gpu.txt 2>&1
echo "== pytest -m gpu"; timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -40 | tee $OUT/pytest_gpu.txt
echo "== smoke"; timeout 300 python __graft_entry__.py smoke 2>&1 | tail -5 | tee $OUT/smoke.txt
echo "== bench"; timeout 900 python bench.py --steps 30 --warmup 5 > $OUT/bench.json 2> $OUT/bench.err; tail -3 $OUT/bench.err; cat $OUT/bench.json
if [ "$2" != "noncu" ]; then
echo "== ncu launch list"
MPB200_NO_GRAPH=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 1 > $OUT/ncu_bench.log 2>&1
echo "== ncu full"
# six main kernels of one warm step (graph capture off so every kernel is a plain launch); read back here with
# scripts/ncu_summary.py / ncu_source.py / ncu_traffic.py
MPB200_NO_GRAPH=1 timeout 900 ncu --set full --clock-control none --import-source on \
    -k regex:'rball_fill|rball_count|classify_columns|edges_free|cell_scatter|points_free' -s 12 -c 6 \
    -o $OUT/prof -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 1 > $OUT/ncu_full.log 2>&1
ls -la $OUT
fi
