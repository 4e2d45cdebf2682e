#!/bin/bash
# first (on GPU 0): new single-GPU tests; then the scaling bench
TAG=${1:-r2s}; shift
OUT=gpurun_out/$TAG; mkdir -p $OUT
echo "== tests"; CUDA_VISIBLE_DEVICES=0 timeout 900 python -m pytest tests/test_gpu_closest.py tests/test_gpu_lq.py tests/test_gpu_mc.py -m gpu -q 2>&1 | tail -12 | tee $OUT/pytest.txt
for n in "$@"; do
  echo "== bench x$n"
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29711 \
      bench.py --gpus $n --steps 30 --warmup 5 > $OUT/bench_$n.json 2> $OUT/bench_$n.err
  tail -2 $OUT/bench_$n.err; python - <<PY
import json
d=json.loads(open('$OUT/bench_$n.json').read().strip().splitlines()[-1])
print('x$n', d['exchange'], 'ms', d['ms_per_step'], 'wall', d['wall_ms_per_step_incl_flush'], 'value', d['value'], 'phase', d['phase_ms'])
print('   strong', d['strong_scaling']); print('   mc', d['mc']['value'], d['mc']['hits'], 'e2e', d['e2e'], 'parity', d['parity_checked'])
PY
done
