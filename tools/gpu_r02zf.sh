#!/bin/bash
# Round-2 GPU call zf: validation of the final HEAD (K1 staging index arithmetic, 2-D grid; e2e as the median of three passes):
# full GPU suite, smoke, bench (both arms), launch list of one step, K1 launch metrics
set -u
O=gpurun_out/r02zf
mkdir -p $O
timeout 600 python -m pytest tests -q -m gpu > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/rc.txt
tail -2 $O/pytest_gpu.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke rc=$?" >> $O/rc.txt
tail -1 $O/smoke.log
timeout 200 python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_ref.json 2> $O/bench_ref.err; echo "bench ref rc=$?" >> $O/rc.txt
timeout 400 python bench.py --steps 20 --warmup 3 > $O/bench.json 2> $O/bench.err; echo "bench rc=$?" >> $O/rc.txt
tail -3 $O/bench.err
timeout 200 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:letterbox --csv --log-file $O/k1_traffic.csv \
    python tools/kernel_bench.py --only k1 --profile > $O/ncu_k1.log 2>&1; echo "ncu k1 metrics rc=$?" >> $O/rc.txt
timeout 300 ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum \
    --clock-control none --csv --log-file $O/launches.csv \
    python bench.py --steps 1 --warmup 3 --no-4k --no-c1 --no-cpu-baseline --profile-region > $O/ncu_launches.log 2>&1; echo "launch list rc=$?" >> $O/rc.txt
cat $O/rc.txt
