#!/bin/bash
# Round-2 GPU call c: K3a/K3c (256 threads, 4 pixels in flight), K3b v2 (coefficient table, packed intermediate, 4-wide
# vertical pass), bucketed trunk batch, process_frame on the device frame
set -u
O=gpurun_out/r02c
mkdir -p $O
timeout 900 python -m pytest tests -x -q -m gpu > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/rc.txt
tail -5 $O/pytest_gpu.log
timeout 300 python tools/kernel_bench.py --only k3 > $O/k3.jsonl 2>&1; echo "k3 rc=$?" >> $O/rc.txt
cat $O/k3.jsonl
timeout 600 python bench.py --steps 10 --warmup 3 > $O/bench.json 2> $O/bench.err; echo "bench rc=$?" >> $O/rc.txt
tail -3 $O/bench.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"jersey_color|color_features|mnv3_prep_fast" -c 3 -f -o $O/k3 \
    python tools/kernel_bench.py --only k3 --profile > $O/ncu_k3.log 2>&1; echo "ncu k3 rc=$?" >> $O/rc.txt
cat $O/rc.txt
