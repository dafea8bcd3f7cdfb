#!/bin/bash
# Round-2 GPU call z: K3a / K3c threads per crop (256 = default, 128, 64: libhvb_k3t*.so), alone; colour parity tests on the 128 build
set -u
O=gpurun_out/r02z
mkdir -p $O
L=$PWD/hockey-vision-analytics_b200/hvb
for t in "" _k3t128 _k3t64; do
  HVB_LIB=$L/libhvb$t.so timeout 120 python tools/kernel_bench.py --only k3 --reps 50 > $O/k3$t.jsonl 2>&1
  echo "variant '$t'"; grep -h '"us"' $O/k3$t.jsonl | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print('   ', d['kernel'], d['us'])"
done
HVB_LIB=$L/libhvb_k3t128.so timeout 400 python -m pytest tests/test_gpu_color.py tests/test_gpu_team_segmentation.py -q -m gpu > $O/pytest_128.log 2>&1; echo "pytest 128 rc=$?" >> $O/rc.txt
tail -2 $O/pytest_128.log
cat $O/rc.txt
