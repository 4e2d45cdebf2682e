#!/bin/bash
# round-2 K2 iteration: table parity (narrow + forced-wide keys), A/B bench old vs cell-centric K2, ncu of the new kernels
TAG=${1:-r2k2}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > $OUT/gpu.txt 2>&1
echo "== parity (new K2)"; timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_edge_cases.py tests/test_gpu_fmt.py -m gpu -x -q 2>&1 | tail -15 | tee $OUT/pytest_new.txt
echo "== parity (wide keys)"; MPB200_FORCE_WIDE=1 timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "k1_k2 or k2_ or full_size or edges_and_points" 2>&1 | tail -8 | tee $OUT/pytest_wide.txt
for v in "MPB200_K2=new" "MPB200_K2=old"; do
  echo "== bench $v"; env $v timeout 600 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --e2e-steps 2 2> $OUT/bench_$v.err | tee $OUT/bench_$v.json | python -c "
import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['phase_ms'], d['roofline']['frac'], d.get('renumbered_samples'))"
  tail -3 $OUT/bench_$v.err
done
if [ "$2" != "noncu" ]; then
echo "== ncu full"
MPB200_NO_GRAPH=1 timeout 900 ncu --set full --clock-control none --import-source on \
    -k regex:'cell_fill2|cell_count2|cell_scatter|cell_histogram' -s 8 -c 4 \
    -o $OUT/prof -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 1 > $OUT/ncu_full.log 2>&1
tail -3 $OUT/ncu_full.log
ls -la $OUT
fi
