#!/bin/bash
# Round-2 GPU call e: probes (4K e2e overlap, process_frame phases)
set -u
O=gpurun_out/r02e
mkdir -p $O
ls /usr/lib/x86_64-linux-gnu | grep -i -E "nvcuvid|nvidia-encode" > $O/nvdec_probe.txt 2>&1; echo "---headers" >> $O/nvdec_probe.txt
find / \( -name "nvcuvid.h" -o -name "cuviddec.h" \) 2>/dev/null | grep -v proc >> $O/nvdec_probe.txt
nvidia-smi topo -m > $O/topo.txt 2>&1; nproc >> $O/topo.txt; lscpu | grep -E "NUMA|Model name|Socket" >> $O/topo.txt
timeout 600 python tools/probe_e2e4k.py > $O/probe.jsonl 2> $O/probe.err; echo "probe rc=$?" >> $O/rc.txt
cat $O/probe.jsonl; tail -3 $O/probe.err; cat $O/nvdec_probe.txt
