#!/bin/bash
# round 2, last call: equal CTA shares of the mini-batch chains (A/B against variant e0), GPU suite, smoke, C2 bench (10 steps),
# launch list of the bench command
cd "${GRAFT_REPO_ROOT:-/root/repo}"
P=$PWD/resilient-consensus-based-marl_b200/rcmarl
for v in "" _e0 ""; do
  echo "== librcmarl$v.so"
  RCMARL_LIB=$P/librcmarl$v.so timeout 300 python tools/prof_mb.py 4096 3000 3 2>&1 | tail -1
done
echo "== pytest -m gpu"
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -4
echo "== smoke"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
echo "== bench C2 (10 steps)"
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r02_bench_c2.json 2> gpurun_out/r02_bench_c2.err; tail -c 300 gpurun_out/r02_bench_c2.json; echo
python - <<'PY'
import json
d = json.load(open("gpurun_out/r02_bench_c2.json"))
print("BENCH", d["value"], d["ms_per_step"], d["gpu_launches"], d["e2e"]["value"], d["roofline"]["frac"], {k: (v.get("ms_per_launch") or v.get("us_per_step")) for k, v in d["roofline"]["regimes"].items()}, d["breakdown_ms"])
PY
echo "== launch list of the bench command"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/r02_launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --no-consensus > /dev/null 2>&1; wc -l gpurun_out/r02_launches.csv
