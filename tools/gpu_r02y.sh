#!/bin/bash
# Round-2 GPU call y: register budget of the two K1 plan kinds with 128-thread blocks (HVB_K1_FLIP_BUDGET=1 swaps
# 64 registers / 8 blocks per SM and 48 registers / 10 blocks per SM), alone and inside the bench step
set -u
O=gpurun_out/r02y
mkdir -p $O
timeout 120 python tools/kernel_bench.py --only k1 --reps 50 > $O/k1_default.jsonl 2>&1
HVB_K1_FLIP_BUDGET=1 timeout 120 python tools/kernel_bench.py --only k1 --reps 50 > $O/k1_flip.jsonl 2>&1
for t in default flip; do python -c "
import json; print('$t', ' '.join('%.1f' % json.loads(l)['us'] for l in open('$O/k1_$t.jsonl') if l.startswith('{')))"; done
for t in default flip default2; do
  f=0; [ "$t" = "flip" ] && f=1
  HVB_K1_FLIP_BUDGET=$f timeout 300 python bench.py --steps 8 --warmup 3 --no-c1 --no-cpu-baseline > $O/bench_$t.json 2> $O/bench_$t.err; echo "bench $t rc=$?" >> $O/rc.txt
  python -c "
import json; d=json.load(open('$O/bench_$t.json')); k=d['roofline_k1a']; r=d['roofline_4k']; print('$t', 'value', round(d['value'],1), 'k1a us', round(1e3*k['avg_launch_ms'],1), 'frac', round(k['frac'],3), 'k1b ms', round(r['avg_launch_ms'],4), 'frac', round(r['frac'],3), d['clocks'])"
done
cat $O/rc.txt
