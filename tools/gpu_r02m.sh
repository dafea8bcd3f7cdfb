#!/bin/bash
# Round-2 GPU call m: upload routes (pinned direct / staged), bench with the pinned-source e2e, K1 traffic capture (x16 K1b)
set -u
O=gpurun_out/r02m
mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_upload.py tests/test_gpu_video.py tests/test_gpu_letterbox.py -q -x -m gpu > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/rc.txt
tail -3 $O/pytest.log
timeout 700 python bench.py --steps 10 --warmup 3 > $O/bench.json 2> $O/bench.err; echo "bench rc=$?" >> $O/rc.txt
tail -3 $O/bench.err
timeout 300 python tools/kernel_bench.py --only k1 > $O/k1.jsonl 2>&1; cat $O/k1.jsonl
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:letterbox --csv --log-file $O/k1_traffic.csv \
    python tools/kernel_bench.py --only k1 --profile > $O/ncu_k1.log 2>&1; echo "ncu k1 rc=$?" >> $O/rc.txt
cat $O/rc.txt
