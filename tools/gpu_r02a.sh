#!/bin/bash
# Round-2 first GPU call: evidence at HEAD (VERDICT "next" #3).  Run under gpurun from the repo root:
#   gpurun --timeout 1500 -- 'bash tools/gpu_r02a.sh'
# Every ncu command runs only after the same command has exited 0 without ncu.
set -u
O=gpurun_out/r02a
mkdir -p $O
nvidia-smi --query-gpu=name,driver_version,clocks.max.sm,clocks.max.mem --format=csv > $O/gpu.txt 2>&1
ls /usr/lib/x86_64-linux-gnu | grep -i -E "nvcuvid|nvidia-encode|libcuda" > $O/nvdec_probe.txt 2>&1
find / -name "nvcuvid.h" -o -name "cuviddec.h" 2>/dev/null | grep -v proc >> $O/nvdec_probe.txt
nproc >> $O/gpu.txt; lscpu | grep -E "NUMA|Model name|Socket" >> $O/gpu.txt

# 1. the untested 128x192 tile variant of K6 (decide: keep or delete)
HVB_PW_TILE=192 timeout 300 python -m pytest tests/test_gpu_pointwise_conv.py -x -q > $O/pytest_pw192.log 2>&1; echo "pw192 rc=$?" >> $O/rc.txt
HVB_PW_TILE=192 timeout 300 python tools/kernel_bench.py --only k6 > $O/k6_tile192.jsonl 2>&1
timeout 300 python tools/kernel_bench.py --only k6,k3 > $O/k6_k3_default.jsonl 2>&1; echo "kb rc=$?" >> $O/rc.txt

# 2. K6 routing A/B through bench.py, same box, 3 runs each, interleaved
for i in 1 2 3; do
  HVB_K6=0 timeout 300 python bench.py --steps 10 --warmup 3 --no-4k --no-cpu-baseline > $O/bench_k6off_$i.json 2>$O/bench_k6off_$i.err
  HVB_K6=1 timeout 300 python bench.py --steps 10 --warmup 3 --no-4k --no-cpu-baseline > $O/bench_k6on_$i.json 2>$O/bench_k6on_$i.err
done
echo "ab done" >> $O/rc.txt

# 3. launch list of one step at HEAD with DRAM bytes
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum \
    --clock-control none --csv --log-file $O/launches.csv \
    python bench.py --steps 1 --warmup 3 --no-4k --no-cpu-baseline --profile-region > $O/ncu_launches.log 2>&1; echo "launch list rc=$?" >> $O/rc.txt

# 4. --set full on K6 and the K3 family
timeout 600 ncu --set full --clock-control none --import-source on -k regex:pointwise_conv -c 8 -f -o $O/k6 \
    python tools/kernel_bench.py --only k6 --profile > $O/ncu_k6.log 2>&1; echo "ncu k6 rc=$?" >> $O/rc.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"jersey_color|color_features|mnv3_prep" -c 6 -f -o $O/k3 \
    python tools/kernel_bench.py --only k3 --profile > $O/ncu_k3.log 2>&1; echo "ncu k3 rc=$?" >> $O/rc.txt
ls -la $O
