#!/bin/bash
# Round-2 GPU call j: K3b (staging with loads in flight, no clamps, half-crop work items) and K3a (prologue, registers for
# 5 / 6 CTAs per SM): parity tests, A/B of the register variants, ncu of the default build
set -u
O=gpurun_out/r02j
mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_color.py tests/test_gpu_mnv3.py tests/test_gpu_team_segmentation.py tests/test_gpu_team_e2e.py -q -x -m gpu > $O/pytest_k3.log 2>&1; echo "pytest rc=$?" >> $O/rc.txt
tail -3 $O/pytest_k3.log
for tag in default k3a4 k3a6; do
  lib=hockey-vision-analytics_b200/hvb/libhvb.so; [ $tag != default ] && lib=hockey-vision-analytics_b200/hvb/libhvb_$tag.so
  for i in 1 2; do HVB_LIB=$PWD/$lib timeout 200 python tools/kernel_bench.py --only k3 --reps 50 > $O/k3_${tag}_$i.jsonl 2>&1; done
  echo "== $tag"; cat $O/k3_${tag}_2.jsonl
done
HVB_LIB=$PWD/hockey-vision-analytics_b200/hvb/libhvb_k3a6.so timeout 300 python -m pytest tests/test_gpu_color.py -q -x -m gpu > $O/pytest_k3a6.log 2>&1; echo "pytest k3a6 rc=$?" >> $O/rc.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"jersey_color|color_features|mnv3_prep_fast" -c 3 -f -o $O/k3 \
    python tools/kernel_bench.py --only k3 --profile > $O/ncu_k3.log 2>&1; echo "ncu k3 rc=$?" >> $O/rc.txt
cat $O/rc.txt
