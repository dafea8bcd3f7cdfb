#!/bin/bash
# Round-2 GPU call zd: validation of the final HEAD — full GPU suite, smoke, bench (both arms), launch list of one step,
# K1 under ncu (launch metrics of every kernel_bench configuration + one --set full capture), K1 / K3 alone
set -u
O=gpurun_out/r02zd
mkdir -p $O
timeout 900 python -m pytest tests -q -m gpu > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/rc.txt
tail -4 $O/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke rc=$?" >> $O/rc.txt
tail -1 $O/smoke.log
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_ref.json 2> $O/bench_ref.err; echo "bench ref rc=$?" >> $O/rc.txt
timeout 700 python bench.py --steps 20 --warmup 3 > $O/bench.json 2> $O/bench.err; echo "bench rc=$?" >> $O/rc.txt
tail -3 $O/bench.err
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum \
    --clock-control none --csv --log-file $O/launches.csv \
    python bench.py --steps 1 --warmup 3 --no-4k --no-c1 --no-cpu-baseline --profile-region > $O/ncu_launches.log 2>&1; echo "launch list rc=$?" >> $O/rc.txt
timeout 200 python tools/kernel_bench.py --only k1,k3 --reps 50 > $O/k1_k3.jsonl 2>&1; cat $O/k1_k3.jsonl
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:letterbox --csv --log-file $O/k1_traffic.csv \
    python tools/kernel_bench.py --only k1 --profile > $O/ncu_k1.log 2>&1; echo "ncu k1 metrics rc=$?" >> $O/rc.txt
timeout 400 ncu --set full --clock-control none --import-source on -k regex:letterbox -c 7 -f -o $O/k1 \
    python tools/kernel_bench.py --only k1 --profile > $O/ncu_k1_full.log 2>&1; echo "ncu k1 full rc=$?" >> $O/rc.txt
ls -la $O
cat $O/rc.txt
