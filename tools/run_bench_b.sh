#!/bin/bash
cd /root/repo; mkdir -p gpurun_out
# launch list of the pile64 step (last 3 of 158 steps: skip the first 155 x ~14 launches)
ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 2300 -c 60 --csv --log-file gpurun_out/r2_launches_pile64.csv python tools/profile_scene.py pile64 4096 155 3 > gpurun_out/prof_pile64.log 2>&1
python tools/launch_summary.py gpurun_out/r2_launches_pile64.csv 20 > gpurun_out/r2_launches_pile64.txt 2>&1
cat gpurun_out/r2_launches_pile64.txt
timeout 300 python tools/gpu_solve6_check.py time pile64 4096 auto
start=$(date +%s)
python bench.py --steps 30 --warmup 5 > gpurun_out/bench_a.json 2> gpurun_out/bench_a.err
echo "bench wall: $(( $(date +%s) - start )) s"; tail -3 gpurun_out/bench_a.err
python - <<'PY'
import json
j = json.loads(open('/root/repo/gpurun_out/bench_a.json').read().strip().splitlines()[-1])
oc = j.pop("other_configs", {})
print(json.dumps(j)[:2500])
for k, v in oc.items():
    print(k, json.dumps(v)[:900])
PY
start=$(date +%s)
python bench.py --impl reference --steps 30 --warmup 5 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
echo "reference arm wall: $(( $(date +%s) - start )) s"; head -c 1200 gpurun_out/bench_ref.json
