#!/bin/bash
# Round-2 GPU call l (8 GPUs): bench at N=8 (weak scaling, drop-in figures of every rank, C-ABI exchange inside bench),
# the C-ABI exchange against torch.distributed on 8 ranks, K3b tap specialisation parity + line
set -u
O=gpurun_out/r02l
mkdir -p $O
nproc > $O/host.txt; nvidia-smi topo -m >> $O/host.txt 2>&1
timeout 200 python -m pytest tests/test_gpu_mnv3.py tests/test_gpu_team_e2e.py -q -x -m gpu > $O/pytest_k3b.log 2>&1; echo "pytest rc=$?" >> $O/rc.txt
tail -2 $O/pytest_k3b.log
timeout 100 python tools/kernel_bench.py --only k3 --reps 50 > $O/k3.jsonl 2>&1; grep K3b $O/k3.jsonl
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29531 tools/check_comm_multigpu.py > $O/comm8.jsonl 2> $O/comm8.err; echo "comm8 rc=$?" >> $O/rc.txt
cat $O/comm8.jsonl
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29532 bench.py --gpus 8 --steps 10 --warmup 3 > $O/bench_n8.json 2> $O/bench_n8.err; echo "bench n8 rc=$?" >> $O/rc.txt
tail -3 $O/bench_n8.err
cat $O/rc.txt
