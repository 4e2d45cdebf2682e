#!/bin/bash
# one bench line at N GPUs (gpurun --gpus N): usage scripts/r2_bench_n.sh N [steps]
N=${1:-2}; K=${2:-30}; OUT=gpurun_out/final_scale; mkdir -p $OUT
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2972$N \
    bench.py --gpus $N --steps $K --warmup 5 > $OUT/bench_$N.json 2> $OUT/bench_$N.err
grep -v "OMP_NUM\|^\*\*\*" $OUT/bench_$N.err | tail -3
python - <<PY
import json
d=json.loads(open('$OUT/bench_$N.json').read().strip().splitlines()[-1])
print('x$N', d['exchange'], 'ms', d['ms_per_step'], 'value', d['value'], 'phase', d['phase_ms'])
print('   strong', d['strong_scaling']['ms_per_step'], 'mc', d['mc']['value'], d['mc']['hits'], 'e2e', d['e2e']['ms_per_step'], 'parity', d['parity_checked'])
PY
