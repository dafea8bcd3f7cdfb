#!/bin/bash
# Round-2 GPU call q: K3a / K3c with the incremental pixel walk — parity + lines; upload / spectral robustness changes
set -u
O=gpurun_out/r02q
mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_color.py tests/test_gpu_team_segmentation.py tests/test_gpu_team_e2e.py tests/test_gpu_upload.py tests/test_gpu_spectral.py tests/test_gpu_video.py -q -x -m gpu > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/rc.txt
tail -3 $O/pytest.log
for i in 1 2; do timeout 200 python tools/kernel_bench.py --only k3 --reps 50 > $O/k3_$i.jsonl 2>&1; done
cat $O/k3_2.jsonl
cat $O/rc.txt
