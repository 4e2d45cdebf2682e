#!/bin/bash
# multi-GPU check (gpurun --gpus N): exchange test, then bench under torchrun
N=${1:-2}; TAG=${2:-r2m$N}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=index,name,clocks.sm,clocks.max.sm --format=csv > $OUT/gpu.txt 2>&1
nvidia-smi topo -m > $OUT/topo.txt 2>&1
echo "== fullsize + c abi"; timeout 1200 python -m pytest tests/test_gpu_fullsize.py tests/test_gpu_c_abi.py -m gpu -q 2>&1 | tail -8 | tee $OUT/pytest_new.txt
echo "== multi-GPU exchange test"; timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -x -q 2>&1 | tail -15 | tee $OUT/pytest_multi.txt
for n in $(seq 2 2 $N) ; do
  [ $n -eq 6 ] && continue
  echo "== bench x$n"
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29711 \
      bench.py --gpus $n --steps 30 --warmup 5 > $OUT/bench_$n.json 2> $OUT/bench_$n.err
  tail -3 $OUT/bench_$n.err; cat $OUT/bench_$n.json
  echo "== bench x$n (nccl exchange)"
  MPB200_EXCHANGE=nccl timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29712 \
      bench.py --gpus $n --steps 30 --warmup 5 --no-secondary --e2e-steps 1 > $OUT/bench_${n}_nccl.json 2> $OUT/bench_${n}_nccl.err
  tail -3 $OUT/bench_${n}_nccl.err; python -c "
import json,sys; d=json.loads(open('$OUT/bench_${n}_nccl.json').read().strip().splitlines()[-1]); print('nccl', d['ms_per_step'], d['phase_ms'], d['strong_scaling'])"
done
