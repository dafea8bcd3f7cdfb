#!/bin/bash
# Round-2 GPU call zb: host-side profile of the drop-in (VideoProcessor.process_video_chunked, chunk 32)
set -u
O=gpurun_out/r02zc
mkdir -p $O
timeout 400 python bench.py --steps 3 --warmup 3 --no-4k --no-c1 --no-cpu-baseline --profile-dropin > $O/bench.json 2> $O/bench.err; echo "bench rc=$?" >> $O/rc.txt
python -c "
import json; d=json.load(open('$O/bench.json')); e=d['extra']; print('value', round(d['value'],1), 'e2e', round(d['e2e']['value'],1), 'dropin', e['clip_chunked_drop_in_fps_chunk32'], e.get('clip_chunked_drop_in_fps_chunk32_runs'), 'frame', e['frame_at_a_time_process_frame_fps_cuda_graph'])"
grep -n "process_frame x20" -A70 $O/bench.err | head -90
cat $O/rc.txt
