#!/bin/bash
# Round-2 GPU call b: the new pieces (K7 device ByteTrack, planted overlay, per-stream work areas, chunk pipeline)
#   gpurun --timeout 1500 -- 'bash tools/gpu_r02b.sh'
set -u
O=gpurun_out/r02b
mkdir -p $O
timeout 900 python -m pytest tests -x -q -m gpu > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/rc.txt
tail -5 $O/pytest_gpu.log
timeout 300 python tools/kernel_bench.py --only track > $O/track.jsonl 2>&1; echo "track rc=$?" >> $O/rc.txt
cat $O/track.jsonl
timeout 600 python bench.py --steps 10 --warmup 3 > $O/bench.json 2> $O/bench.err; echo "bench rc=$?" >> $O/rc.txt
tail -3 $O/bench.err
HVB_K6=0 timeout 600 python bench.py --steps 10 --warmup 3 --no-4k --no-c1 --no-cpu-baseline > $O/bench_k6off.json 2> $O/bench_k6off.err; echo "bench k6off rc=$?" >> $O/rc.txt
timeout 600 python bench.py --steps 10 --warmup 3 --no-4k --no-c1 --no-cpu-baseline --tracker host > $O/bench_hosttracker.json 2> $O/bench_hosttracker.err; echo "bench host tracker rc=$?" >> $O/rc.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke rc=$?" >> $O/rc.txt
cat $O/rc.txt
