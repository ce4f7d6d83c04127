#!/bin/bash
# round 2: validation of the final tree (interleaved kernel removed): GPU suite, smoke, short bench
cd "${GRAFT_REPO_ROOT:-/root/repo}"
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -3
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-consensus > gpurun_out/bench_call28.json 2> gpurun_out/bench_call28.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/bench_call28.json"))
print("BENCH", d["value"], d["ms_per_step"], d["gpu_launches"], d["e2e"]["value"])
PY
timeout 200 python bench.py --impl reference --steps 1 --warmup 0 2>/dev/null | tail -c 300
