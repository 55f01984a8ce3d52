#!/bin/bash
# the round-end sequence in one call: GPU test suite, smoke, bench (own arm)
cd /root/repo; mkdir -p gpurun_out
start=$(date +%s)
timeout 2400 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$? in $(( $(date +%s) - start )) s"; tail -4 gpurun_out/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
start=$(date +%s)
python bench.py --steps 30 --warmup 5 > gpurun_out/bench_a.json 2> gpurun_out/bench_a.err
echo "bench: $(( $(date +%s) - start )) s"; tail -3 gpurun_out/bench_a.err
python - <<'PY'
import json
j = json.loads(open('/root/repo/gpurun_out/bench_a.json').read().strip().splitlines()[-1])
oc = j.pop("other_configs", {})
print("headline", j["value"], j["ms_per_step"], j["e2e"]["value"], j["roofline"]["frac"])
for k, v in oc.items():
    print(k, v.get("ms_per_step"), v.get("body_steps_per_sec"), json.dumps(v.get("roofline"))[:260], json.dumps(v.get("cpu_baseline"))[:200])
PY
