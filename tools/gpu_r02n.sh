#!/bin/bash
# Round-2 GPU call n (8 GPUs): bench at N=8 with the pinned-source e2e route and clock sampling started early
set -u
O=gpurun_out/r02n
mkdir -p $O
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29532 bench.py --gpus 8 --steps 10 --warmup 3 > $O/bench_n8.json 2> $O/bench_n8.err; echo "bench n8 rc=$?" >> $O/rc.txt
tail -3 $O/bench_n8.err
cat $O/rc.txt
