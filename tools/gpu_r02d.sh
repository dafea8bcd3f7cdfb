#!/bin/bash
# Round-2 GPU call d (re-entry): state of HEAD 4edaae8 — GPU tests, K3/K7 kernel lines, the bench line, smoke, ncu K3
#   gpurun --timeout 1800 -- 'bash tools/gpu_r02d.sh'
set -u
O=gpurun_out/r02d
mkdir -p $O
timeout 900 python -m pytest tests -q -m gpu > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/rc.txt
tail -15 $O/pytest_gpu.log
timeout 300 python tools/kernel_bench.py --only k3,track > $O/k3_track.jsonl 2>&1; echo "kb rc=$?" >> $O/rc.txt
cat $O/k3_track.jsonl
timeout 700 python bench.py --steps 10 --warmup 3 > $O/bench.json 2> $O/bench.err; echo "bench rc=$?" >> $O/rc.txt
tail -5 $O/bench.err
timeout 400 python bench.py --steps 10 --warmup 3 --no-4k --no-c1 --no-cpu-baseline --tracker host > $O/bench_hosttracker.json 2> $O/bench_hosttracker.err; echo "bench host tracker rc=$?" >> $O/rc.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke rc=$?" >> $O/rc.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"jersey_color|color_features|mnv3_prep_fast" -c 3 -f -o $O/k3 \
    python tools/kernel_bench.py --only k3 --profile > $O/ncu_k3.log 2>&1; echo "ncu k3 rc=$?" >> $O/rc.txt
cat $O/rc.txt
