#!/bin/bash
# 2-GPU visit: usage: gpurun --gpus 2 --timeout 900 -- 'bash tools/ddp2.sh [N]'
N=${1:-2}
O=gpurun_out/ddp$N; mkdir -p $O
export NCCL_DEBUG=WARN
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29611 tools/ddp_timeline.py > $O/timeline.txt 2>&1
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29612 bench.py --gpus $N --steps 20 --warmup 5 > $O/bench_train.json 2> $O/bench_train.err
timeout 400 python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu-baseline --no-onbox > $O/bench_train_1gpu.json 2>> $O/bench_train.err
tail -8 $O/timeline.txt; tail -1 $O/bench_train.json | cut -c1-700; tail -1 $O/bench_train_1gpu.json | cut -c1-300; tail -3 $O/bench_train.err
