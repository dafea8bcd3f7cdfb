#!/bin/bash
# Round-2 GPU call o: K1a with two output columns per thread (HVB_K1_COLS=2) against the one-column kernel
set -u
O=gpurun_out/r02o
mkdir -p $O
HVB_K1_COLS=2 timeout 600 python -m pytest tests/test_gpu_letterbox.py tests/test_gpu_fullsize_properties.py -q -x -m gpu > $O/pytest_cols2.log 2>&1; echo "pytest cols2 rc=$?" >> $O/rc.txt
tail -3 $O/pytest_cols2.log
timeout 300 python -m pytest tests/test_gpu_letterbox.py -q -x -m gpu > $O/pytest_cols1.log 2>&1; echo "pytest cols1 rc=$?" >> $O/rc.txt
for i in 1 2; do
  HVB_K1_COLS=1 timeout 200 python tools/kernel_bench.py --only k1 > $O/k1_cols1_$i.jsonl 2>&1
  HVB_K1_COLS=2 timeout 200 python tools/kernel_bench.py --only k1 > $O/k1_cols2_$i.jsonl 2>&1
done
echo "== cols1"; grep K1a $O/k1_cols1_2.jsonl; echo "== cols2"; grep K1a $O/k1_cols2_2.jsonl
for i in 1 2; do
  HVB_K1_COLS=1 timeout 400 python bench.py --steps 10 --warmup 3 --no-4k --no-c1 --no-cpu-baseline > $O/bench_cols1_$i.json 2> $O/bench_cols1_$i.err
  HVB_K1_COLS=2 timeout 400 python bench.py --steps 10 --warmup 3 --no-4k --no-c1 --no-cpu-baseline > $O/bench_cols2_$i.json 2> $O/bench_cols2_$i.err
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r02o/bench_cols*.json')):
    d=json.loads(open(f).read().strip().splitlines()[-1]); k=d['extra']['roofline_k1a']
    print(f.split('/')[-1], round(d['value'],1), round(k['frac'],3), round(k['avg_launch_ms']*1e3,1), d['clocks']['sm_mhz'])
PY
cat $O/rc.txt
