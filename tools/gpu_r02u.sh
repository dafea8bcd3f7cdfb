#!/bin/bash
# Round-2 GPU call u: K1 column loop with the ALU-pipe work moved to the FMA pipe (A/B against the previous kernel
# built as libhvb_k1old.so): letterbox / slicer parity tests, K1 alone, K1a / K1b inside the bench step
set -u
O=gpurun_out/r02u
mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_letterbox.py tests/test_gpu_fullsize_properties.py tests/test_gpu_merge.py -q -m gpu > $O/pytest_k1.log 2>&1; echo "pytest rc=$?" >> $O/rc.txt
tail -3 $O/pytest_k1.log
timeout 200 python tools/kernel_bench.py --only k1 --reps 50 > $O/k1_new.jsonl 2>&1; cat $O/k1_new.jsonl
HVB_LIB=$PWD/hockey-vision-analytics_b200/hvb/libhvb_k1old.so timeout 200 python tools/kernel_bench.py --only k1 --reps 50 > $O/k1_old.jsonl 2>&1; cat $O/k1_old.jsonl
timeout 400 python bench.py --steps 10 --warmup 3 --no-c1 --no-cpu-baseline > $O/bench_new.json 2> $O/bench_new.err; echo "bench new rc=$?" >> $O/rc.txt
HVB_LIB=$PWD/hockey-vision-analytics_b200/hvb/libhvb_k1old.so timeout 400 python bench.py --steps 10 --warmup 3 --no-c1 --no-cpu-baseline > $O/bench_old.json 2> $O/bench_old.err; echo "bench old rc=$?" >> $O/rc.txt
python - <<'PY'
import json
for t in ("new", "old"):
    try:
        d = json.load(open("gpurun_out/r02u/bench_%s.json" % t))
        print(t, "value", round(d["value"], 1), "e2e", round(d["e2e"]["value"], 1), "k1a", d["extra"].get("roofline_k1a"), "k1b frac", d["roofline_4k"]["frac"], d["roofline_4k"]["avg_launch_ms"], d["clocks"])
    except Exception as e:
        print(t, "failed", e)
PY
cat $O/rc.txt
