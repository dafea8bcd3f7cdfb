#!/bin/bash
# Round-2 GPU call v: K1 column-loop variants (libhvb_k1*.so, see the HVB_K1_* switches in csrc/k1_letterbox.cu) alone,
# then the two fastest and the previous kernel inside the bench step
set -u
O=gpurun_out/r02w
mkdir -p $O
L=$PWD/hockey-vision-analytics_b200/hvb
timeout 300 python -m pytest tests/test_gpu_letterbox.py -q -m gpu -x > $O/pytest_k1.log 2>&1; echo "pytest rc=$?" >> $O/rc.txt
tail -2 $O/pytest_k1.log
for t in "" _k1b _k1c _k1d _k1e _k1f _k1old; do
  HVB_LIB=$L/libhvb$t.so timeout 120 python tools/kernel_bench.py --only k1 --reps 50 > $O/k1$t.jsonl 2>&1
done
python - <<'PY' > gpurun_out/r02w/pick.txt
import json, glob, os
res = {}
for f in sorted(glob.glob("gpurun_out/r02w/k1*.jsonl")):
    tag = os.path.basename(f)[2:-6] or "_new"
    rows = [json.loads(l) for l in open(f) if l.startswith("{")]
    res[tag] = {r["kernel"]: r["us"] for r in rows}
    print(tag, " ".join("%.1f" % r["us"] for r in rows))
key = "K1a 1080p->736x1280 x64"
order = sorted((t for t in res if t != "_k1old" and key in res[t]), key=lambda t: res[t][key])
print("PICK", " ".join(order[:2]))
PY
cat $O/pick.txt
for t in $(grep PICK $O/pick.txt | cut -d' ' -f2-) _k1old; do
  s=$t; [ "$t" = "_new" ] && s=""
  HVB_LIB=$L/libhvb$s.so timeout 300 python bench.py --steps 8 --warmup 3 --no-4k --no-c1 --no-cpu-baseline > $O/bench$t.json 2> $O/bench$t.err; echo "bench $t rc=$?" >> $O/rc.txt
  python -c "
import json; d=json.load(open('$O/bench$t.json')); k=d['extra']['roofline_k1a']; print('$t', 'value', round(d['value'],1), 'k1a us', round(1e3*k['avg_launch_ms'],1), 'frac', round(k['frac'],3), d['clocks'])"
done
cat $O/rc.txt
