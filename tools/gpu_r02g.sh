#!/bin/bash
# Round-2 GPU call g: K8 (device spectral clustering kernels), the C ABI's feature exchange on one rank, trunk graph fix
set -u
O=gpurun_out/r02g
mkdir -p $O
timeout 900 python -m pytest tests -q -x -m gpu > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/rc.txt
tail -25 $O/pytest_gpu.log
timeout 300 python -m pytest tests/test_gpu_spectral.py -q -s -m gpu > $O/pytest_spectral.log 2>&1
grep -E "N=|fit" $O/pytest_spectral.log
timeout 300 python tools/kernel_bench.py --only k8 > $O/k8.jsonl 2>&1; echo "k8 rc=$?" >> $O/rc.txt
cat $O/k8.jsonl
timeout 700 python bench.py --steps 10 --warmup 3 > $O/bench.json 2> $O/bench.err; echo "bench rc=$?" >> $O/rc.txt
tail -5 $O/bench.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_ref.json 2> $O/bench_ref.err; echo "bench ref rc=$?" >> $O/rc.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke rc=$?" >> $O/rc.txt
cat $O/rc.txt
