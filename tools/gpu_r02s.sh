#!/bin/bash
# Round-2 GPU call s: K3a / K3c with explicit shared-window addresses (tables, histogram counters), 64-register K3a
set -u
O=gpurun_out/r02s
mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_color.py tests/test_gpu_team_segmentation.py tests/test_gpu_team_e2e.py -q -x -m gpu > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/rc.txt
tail -3 $O/pytest.log
for i in 1 2 3; do timeout 200 python tools/kernel_bench.py --only k3 --reps 50 > $O/k3_$i.jsonl 2>&1; done
grep -h "K3a\|K3c" $O/k3_*.jsonl
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"jersey_color|color_features|mnv3_prep_fast" -c 3 -f -o $O/k3 \
    python tools/kernel_bench.py --only k3 --profile > $O/ncu_k3.log 2>&1; echo "ncu k3 rc=$?" >> $O/rc.txt
cat $O/rc.txt
