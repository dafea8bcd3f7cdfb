#!/bin/bash
# Round-2 GPU call k: full GPU suite at HEAD, K3 lines, K7 inlined vs outlined building blocks, bench
set -u
O=gpurun_out/r02k
mkdir -p $O
timeout 900 python -m pytest tests -q -x -m gpu > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/rc.txt
tail -4 $O/pytest_gpu.log
timeout 200 python tools/kernel_bench.py --only k3 --reps 50 > $O/k3.jsonl 2>&1; cat $O/k3.jsonl
for i in 1 2; do
  timeout 200 python tools/kernel_bench.py --only track > $O/track_inline_$i.jsonl 2>&1
  HVB_LIB=$PWD/hockey-vision-analytics_b200/hvb/libhvb_k7o.so timeout 200 python tools/kernel_bench.py --only track > $O/track_outline_$i.jsonl 2>&1
done
echo "== inline"; grep K7 $O/track_inline_2.jsonl; echo "== outline"; grep K7 $O/track_outline_2.jsonl
HVB_LIB=$PWD/hockey-vision-analytics_b200/hvb/libhvb_k7o.so timeout 300 python -m pytest tests/test_gpu_bytetrack.py tests/test_gpu_video.py -q -x -m gpu > $O/pytest_k7o.log 2>&1; echo "pytest k7o rc=$?" >> $O/rc.txt
timeout 700 python bench.py --steps 10 --warmup 3 > $O/bench.json 2> $O/bench.err; echo "bench rc=$?" >> $O/rc.txt
tail -3 $O/bench.err
HVB_LIB=$PWD/hockey-vision-analytics_b200/hvb/libhvb_k7o.so timeout 400 python bench.py --steps 10 --warmup 3 --no-4k --no-c1 --no-cpu-baseline > $O/bench_k7o.json 2> $O/bench_k7o.err; echo "bench k7o rc=$?" >> $O/rc.txt
cat $O/rc.txt
