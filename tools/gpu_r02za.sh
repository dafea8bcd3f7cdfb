#!/bin/bash
# Round-2 GPU call za: row unrolling of the K1 column loop (libhvb_u{2,8,16}.so against the default 4), alone and in the step
set -u
O=gpurun_out/r02za
mkdir -p $O
L=$PWD/hockey-vision-analytics_b200/hvb
for t in "" _u2 _u8 _u16; do
  HVB_LIB=$L/libhvb$t.so timeout 120 python tools/kernel_bench.py --only k1 --reps 50 > $O/k1$t.jsonl 2>&1
done
python - <<'PY' > gpurun_out/r02za/pick.txt
import json, glob, os
res = {}
for f in sorted(glob.glob("gpurun_out/r02za/k1*.jsonl")):
    tag = os.path.basename(f)[2:-6] or "_default"
    rows = [json.loads(l) for l in open(f) if l.startswith("{")]
    res[tag] = {r["kernel"]: r["us"] for r in rows}
    print(tag, " ".join("%.1f" % r["us"] for r in rows))
key = "K1a 1080p->736x1280 x64"
order = sorted((t for t in res if t != "_default" and key in res[t]), key=lambda t: res[t][key])
print("PICK", " ".join(order[:1]))
PY
cat $O/pick.txt
for t in $(grep PICK $O/pick.txt | cut -d' ' -f2-) _default; do
  s=$t; [ "$t" = "_default" ] && s=""
  HVB_LIB=$L/libhvb$s.so timeout 300 python bench.py --steps 8 --warmup 3 --no-c1 --no-cpu-baseline > $O/bench$t.json 2> $O/bench$t.err; echo "bench $t rc=$?" >> $O/rc.txt
  python -c "
import json; d=json.load(open('$O/bench$t.json')); k=d['roofline_k1a']; r=d['roofline_4k']; print('$t', 'value', round(d['value'],1), 'k1a us', round(1e3*k['avg_launch_ms'],1), 'frac', round(k['frac'],3), 'k1b ms', round(r['avg_launch_ms'],4), 'frac', round(r['frac'],3), d['clocks'])"
done
cat $O/rc.txt
