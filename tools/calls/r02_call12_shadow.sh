#!/bin/bash
# round 2: producer bookkeeping moved into the MMA shadows -- tests, timeline, A/B against the previous build on the same box
cd "${GRAFT_REPO_ROOT:-/root/repo}"
P=$PWD/resilient-consensus-based-marl_b200/rcmarl
echo "== pytest -m gpu"
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -4
echo "== timeline"
RCMARL_LIB=$P/librcmarl_tl.so timeout 200 python tools/ws_timeline.py 2>&1 | tail -12
for v in "" _prev ""; do
  echo "== librcmarl$v.so"
  RCMARL_LIB=$P/librcmarl$v.so timeout 200 python tools/prof_grad.py 12288000 8 5 2>&1 | tail -1
  RCMARL_LIB=$P/librcmarl$v.so timeout 300 python tools/prof_mb.py 4096 3000 3 2>&1 | tail -1
done
echo "== bench (short)"
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_call12.json 2> gpurun_out/bench_call12.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/bench_call12.json"))
print("BENCH", d["value"], d["ms_per_step"], d["gpu_launches"], d["e2e"]["value"], json.dumps(d["roofline"])[:600])
PY
