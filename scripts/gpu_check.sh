#!/bin/bash
# One gpurun call: GPU tests, smoke, bench, ncu launch list + full capture of the top kernels.
# usage: scripts/gpu_check.sh [tag] [noncu]
TAG=${1:-run}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > $OUT/gpu.txt 2>&1
echo "== pytest -m gpu"; timeout 1800 python -m pytest tests -m gpu -q 2>&1 | tail -12 | tee $OUT/pytest_gpu.txt
echo "== smoke"; timeout 300 python __graft_entry__.py smoke 2>&1 | tail -5 | tee $OUT/smoke.txt
echo "== bench"; timeout 1200 python bench.py --steps 50 --warmup 5 > $OUT/bench.json 2> $OUT/bench.err; tail -3 $OUT/bench.err; cut -c1-1500 $OUT/bench.json
echo "== reference arm"; timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $OUT/bench_reference.json 2> $OUT/bench_reference.err; cut -c1-900 $OUT/bench_reference.json
if [ "$2" != "noncu" ]; then
echo "== ncu launch list"
MPB200_NO_GRAPH=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-secondary --e2e-steps 1 > $OUT/ncu_bench.log 2>&1
echo "== ncu full"
# main kernels of one warm step (graph capture off so every kernel is a plain launch); read back here with
# scripts/ncu_summary.py / ncu_source.py / ncu_traffic.py
MPB200_NO_GRAPH=1 timeout 900 ncu --set full --clock-control none --import-source on \
    -k regex:'rball_fill|rball_count|classify_columns|edges_free|cell_scatter|points_free|cell_histogram' -s 14 -c 7 \
    -o $OUT/prof -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-secondary --e2e-steps 1 > $OUT/ncu_full.log 2>&1
tail -2 $OUT/ncu_full.log
ls -la $OUT
fi
