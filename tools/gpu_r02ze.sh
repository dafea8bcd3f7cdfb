#!/bin/bash
# Round-2 GPU call ze: K1 staging index arithmetic (descriptor register, incremental (row, group), 2-D grid): parity, alone, in the step
set -u
O=gpurun_out/r02ze
mkdir -p $O
timeout 300 python -m pytest tests/test_gpu_letterbox.py tests/test_gpu_merge.py tests/test_gpu_fullsize_properties.py -q -m gpu > $O/pytest_k1.log 2>&1; echo "pytest rc=$?" >> $O/rc.txt
tail -2 $O/pytest_k1.log
timeout 120 python tools/kernel_bench.py --only k1 --reps 50 > $O/k1.jsonl 2>&1
python -c "
import json; print(' '.join('%.1f' % json.loads(l)['us'] for l in open('$O/k1.jsonl') if l.startswith('{')))"
timeout 300 python bench.py --steps 8 --warmup 3 --no-c1 --no-cpu-baseline > $O/bench.json 2> $O/bench.err; echo "bench rc=$?" >> $O/rc.txt
python -c "
import json; d=json.load(open('$O/bench.json')); k=d['roofline_k1a']; r=d['roofline_4k']; print('value', round(d['value'],1), 'e2e', round(d['e2e']['value'],1), 'k1a us', round(1e3*k['avg_launch_ms'],1), 'frac', round(k['frac'],3), 'k1b ms', round(r['avg_launch_ms'],4), 'frac', round(r['frac'],3), d['clocks'])"
cat $O/rc.txt
