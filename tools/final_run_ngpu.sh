#!/bin/bash
# Round-end style multi-GPU run on N GPUs of one box: (N = 2: the multi-GPU test suite first,) then the bench line.
set -u
N=$1
O=gpurun_out/final
mkdir -p $O
if [ "$N" = 2 ]; then timeout 600 python -m pytest tests/test_multi_gpu.py -x -q -m gpu > $O/pytest_multi_gpu_2gpus.log 2>&1; tail -2 $O/pytest_multi_gpu_2gpus.log; fi
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29520 + N)) bench.py --gpus $N --steps 20 --warmup 5 > $O/bench_${N}gpu.json 2> $O/bench_${N}gpu.err
tail -c 400 $O/bench_${N}gpu.json
