#!/bin/bash
# Round-2 GPU call h (2 GPUs): the C ABI's feature exchange against torch.distributed on 2 ranks, bench at N=2 (drop-in
# figures of every rank), K8 after the matvec / degree rework
set -u
O=gpurun_out/r02h
mkdir -p $O
nvidia-smi topo -m > $O/topo.txt 2>&1; nproc >> $O/topo.txt
timeout 300 python -m pytest tests/test_gpu_spectral.py tests/test_gpu_comm.py -q -x -m gpu > $O/pytest_k8_comm.log 2>&1; echo "pytest rc=$?" >> $O/rc.txt
tail -3 $O/pytest_k8_comm.log
timeout 300 python tools/kernel_bench.py --only k8 > $O/k8.jsonl 2>&1; echo "k8 rc=$?" >> $O/rc.txt
cat $O/k8.jsonl
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 tools/check_comm_multigpu.py > $O/comm2.jsonl 2> $O/comm2.err; echo "comm2 rc=$?" >> $O/rc.txt
cat $O/comm2.jsonl; tail -3 $O/comm2.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29532 bench.py --gpus 2 --steps 10 --warmup 3 > $O/bench_n2.json 2> $O/bench_n2.err; echo "bench n2 rc=$?" >> $O/rc.txt
tail -3 $O/bench_n2.err
cat $O/rc.txt
