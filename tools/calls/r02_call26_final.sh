#!/bin/bash
# round 2, last call: GPU suite, smoke, C2 bench (10 steps), launch list of the bench command, clip-mean microbench
cd "${GRAFT_REPO_ROOT:-/root/repo}"
echo "== pytest -m gpu"
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -4
echo "== smoke"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
echo "== clip-mean microbench"
timeout 200 python tools/prof_clip_mean.py 0 1 2 4 2>&1 | tail -4
echo "== bench C2 (10 steps)"
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r02_bench_c2.json 2> gpurun_out/r02_bench_c2.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/r02_bench_c2.json"))
print("BENCH", d["value"], d["ms_per_step"], d["gpu_launches"], d["e2e"]["value"], d["roofline"]["frac"], {k: (v.get("ms_per_launch") or v.get("us_per_step")) for k, v in d["roofline"]["regimes"].items()}, {k: round(v["frac"], 3) for k, v in d["consensus_roofline"].items() if isinstance(v, dict)}, d["breakdown_ms"])
PY
echo "== launch list of the bench command"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/r02_launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --no-consensus > /dev/null 2>&1; wc -l gpurun_out/r02_launches.csv
