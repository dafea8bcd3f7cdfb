#!/bin/bash
# Round-2 GPU call r: K3a with shared-memory private histogram counters (default) against the packed register counters
set -u
O=gpurun_out/r02r
mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_color.py tests/test_gpu_team_e2e.py -q -x -m gpu > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/rc.txt
tail -3 $O/pytest.log
for i in 1 2 3; do
  timeout 200 python tools/kernel_bench.py --only k3 --reps 50 > $O/k3_smem_$i.jsonl 2>&1
  HVB_LIB=$PWD/hockey-vision-analytics_b200/hvb/libhvb_k3areg.so timeout 200 python tools/kernel_bench.py --only k3 --reps 50 > $O/k3_reg_$i.jsonl 2>&1
done
echo "== smem hist"; grep -h K3a $O/k3_smem_*.jsonl; echo "== register hist"; grep -h K3a $O/k3_reg_*.jsonl
cat $O/rc.txt
