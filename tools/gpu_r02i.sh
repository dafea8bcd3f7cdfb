#!/bin/bash
# Round-2 GPU call i (2 GPUs): host-side staging through hvb_stage_frames — probe, bench at N=1 and N=2
set -u
O=gpurun_out/r02i
mkdir -p $O
timeout 300 python tools/probe_staging.py > $O/staging.jsonl 2> $O/staging.err; echo "staging rc=$?" >> $O/rc.txt
cat $O/staging.jsonl
timeout 300 python -m pytest tests/test_gpu_video.py tests/test_gpu_merge.py -q -x -m gpu > $O/pytest_video.log 2>&1; echo "pytest rc=$?" >> $O/rc.txt
tail -3 $O/pytest_video.log
timeout 700 python bench.py --steps 10 --warmup 3 > $O/bench_n1.json 2> $O/bench_n1.err; echo "bench n1 rc=$?" >> $O/rc.txt
tail -3 $O/bench_n1.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29532 bench.py --gpus 2 --steps 10 --warmup 3 > $O/bench_n2.json 2> $O/bench_n2.err; echo "bench n2 rc=$?" >> $O/rc.txt
tail -3 $O/bench_n2.err
cat $O/rc.txt
