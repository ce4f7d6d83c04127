#!/bin/bash
# round 2: C3 (10x10 grid, 16 agents, H = 2, 8192 environments) at HEAD, 3 steps
cd "${GRAFT_REPO_ROOT:-/root/repo}"
timeout 900 python bench.py --workload C3 --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --no-consensus > gpurun_out/r02_bench_c3.json 2> gpurun_out/r02_bench_c3.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/r02_bench_c3.json"))
print("BENCH C3", d["value"], d["ms_per_step"], d["gpu_launches"], d["roofline"]["frac"], d["breakdown_ms"])
PY
tail -2 gpurun_out/r02_bench_c3.err
