#!/bin/bash
# scaling run on one multi-GPU box: bench under torchrun for the rank counts given, peer-store exchange
# (and the NCCL form for comparison when the second argument is "nccl")
TAG=${1:-r2s}; CMP=${2:-}; shift; shift
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=index,name,clocks.sm,clocks.max.sm --format=csv > $OUT/gpu.txt 2>&1
for n in "$@"; do
  echo "== bench x$n"
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29711 \
      bench.py --gpus $n --steps 30 --warmup 5 > $OUT/bench_$n.json 2> $OUT/bench_$n.err
  tail -2 $OUT/bench_$n.err; python - <<PY
import json
d=json.loads(open('$OUT/bench_$n.json').read().strip().splitlines()[-1])
print('x$n', d['exchange'], 'ms', d['ms_per_step'], 'value', d['value'], 'phase', d['phase_ms'])
print('   strong', d['strong_scaling']); print('   mc', d['mc']['value'], d['mc']['hits'], 'e2e', d['e2e'], 'parity', d['parity_checked'])
PY
  if [ "$CMP" = "nccl" ]; then
    MPB200_EXCHANGE=nccl timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29712 \
        bench.py --gpus $n --steps 30 --warmup 5 --no-secondary --e2e-steps 1 > $OUT/bench_${n}_nccl.json 2> $OUT/bench_${n}_nccl.err
    python -c "
import json; d=json.loads(open('$OUT/bench_${n}_nccl.json').read().strip().splitlines()[-1]); print('   nccl x$n ms', d['ms_per_step'], 'strong', d['strong_scaling']['ms_per_step'])"
  fi
done
