#!/bin/bash
cd /root/repo
for l in "" /root/repo/ode_b200/variants/r1 ""; do python tools/gpu_ab_time.py stack 4096 256 $l | tail -1; done
timeout 900 python -m pytest tests -m gpu -x -q -k "hybrid_mixed or solver_kernels_bit_exact or device_side" 2>&1 | tail -3
start=$(date +%s)
python bench.py --steps 30 --warmup 5 > gpurun_out/bench_b.json 2> gpurun_out/bench_b.err
echo "bench wall: $(( $(date +%s) - start )) s"; tail -3 gpurun_out/bench_b.err
python - <<'PY'
import json
j = json.loads(open('/root/repo/gpurun_out/bench_b.json').read().strip().splitlines()[-1])
oc = j.pop("other_configs", {})
print("headline", j["value"], j["ms_per_step"], j["e2e"]["value"], j["roofline"]["frac"])
for k, v in oc.items():
    print(k, v.get("ms_per_step"), v.get("body_steps_per_sec"), json.dumps(v.get("cpu_baseline"))[:300], json.dumps(v.get("cpu_threaded_stepper"))[:400] if "cpu_threaded_stepper" in v else "")
PY
