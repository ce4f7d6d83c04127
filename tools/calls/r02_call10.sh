#!/bin/bash
# round 2, GPU call 10: two MMA round trips per tile (L3 of a tile + L1 of the stream's next tile in one batch), straight-line
# consumer steps; full suite, timeline, timings, mini-batch chain, bench
cd "${GRAFT_REPO_ROOT:-/root/repo}"
P=$PWD/resilient-consensus-based-marl_b200/rcmarl
mkdir -p gpurun_out
echo "== pytest -m gpu"
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -8
echo "== producer timeline"
RCMARL_LIB=$P/librcmarl_tl.so timeout 200 python tools/ws_timeline.py 2>&1 | tail -14
echo "== grad timing"
timeout 200 python tools/ab_grad.py time 2>&1 | grep TIMING
echo "== mini-batch chain: exclusive shares (default), interleaved"
timeout 200 python tools/prof_mb.py 4096 3000 3 2>&1 | tail -1
RCMARL_MB_INTERLEAVE=1 timeout 200 python tools/prof_mb.py 4096 3000 3 2>&1 | tail -1
echo "== bench (short)"
timeout 400 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_call10.json 2> gpurun_out/bench_call10.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/bench_call10.json"))
print("BENCH", d["value"], d["ms_per_step"], d["gpu_launches"], d["e2e"]["value"], json.dumps(d["roofline"]["regimes"]), json.dumps(d["breakdown_ms"]))
print(json.dumps({k: round(v["frac"], 3) for k, v in d["consensus_roofline"].items() if isinstance(v, dict)}))
PY
tail -2 gpurun_out/bench_call10.err
