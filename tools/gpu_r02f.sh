#!/bin/bash
# Round-2 GPU call f: persistent staging state of the process_stream pipelines, trunk forward as a CUDA graph per bucket,
# one D2H in predict; then the evidence at this HEAD: launch list of one step (with K7 and planted detections), ncu K7 + K2a
set -u
O=gpurun_out/r02f
mkdir -p $O
timeout 900 python -m pytest tests -q -x -m gpu > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/rc.txt
tail -5 $O/pytest_gpu.log
timeout 700 python bench.py --steps 10 --warmup 3 > $O/bench.json 2> $O/bench.err; echo "bench rc=$?" >> $O/rc.txt
tail -5 $O/bench.err
timeout 300 python tools/probe_e2e4k.py > $O/probe.jsonl 2> $O/probe.err; echo "probe rc=$?" >> $O/rc.txt
cat $O/probe.jsonl
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum \
    --clock-control none --csv --log-file $O/launches.csv \
    python bench.py --steps 1 --warmup 3 --no-4k --no-c1 --no-cpu-baseline --profile-region > $O/ncu_launches.log 2>&1; echo "launch list rc=$?" >> $O/rc.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"bytetrack_kernel|decode_nms_kernel" -c 4 -f -o $O/k7_k2a \
    --profile-from-start off python bench.py --steps 1 --warmup 3 --no-4k --no-c1 --no-cpu-baseline --profile-region > $O/ncu_k7.log 2>&1; echo "ncu k7 rc=$?" >> $O/rc.txt
cat $O/rc.txt
