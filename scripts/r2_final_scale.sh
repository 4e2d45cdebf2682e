#!/bin/bash
# final scaling capture: multi-GPU exchange test, then bench at 8, 4, 2, 1 on one box
TAG=${1:-z_scale}; OUT=gpurun_out/$TAG; mkdir -p $OUT
echo "== multi-GPU exchange test"; timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -x -q 2>&1 | tail -4 | tee $OUT/pytest_multi.txt
for n in 8 4 2; do
  echo "== bench x$n"
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2971$n \
      bench.py --gpus $n --steps 40 --warmup 5 > $OUT/bench_$n.json 2> $OUT/bench_$n.err
  grep -v "OMP_NUM\|^\*\*\*" $OUT/bench_$n.err | tail -2; python - <<PY
import json
d=json.loads(open('$OUT/bench_$n.json').read().strip().splitlines()[-1])
print('x$n', d['exchange'], 'ms', d['ms_per_step'], 'wall', d['wall_ms_per_step_incl_flush'], 'value', d['value'], 'phase', d['phase_ms'])
print('   per_rank', [(r['ms_per_step'], r['exchange_incl_wait']) for r in d['per_rank']])
print('   strong', d['strong_scaling']['ms_per_step'], 'mc', d['mc']['value'], d['mc']['hits'], 'e2e', d['e2e']['ms_per_step'], 'parity', d['parity_checked'])
PY
done
echo "== bench x1"; timeout 600 python bench.py --steps 40 --warmup 5 --no-secondary --no-cpu-baseline > $OUT/bench_1.json 2> $OUT/bench_1.err
python -c "
import json; d=json.loads(open('$OUT/bench_1.json').read().strip().splitlines()[-1]); print('x1 ms', d['ms_per_step'], 'value', d['value'], d['phase_ms'])"
