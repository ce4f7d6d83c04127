#!/bin/bash
# round 2, GPU call 9: default library = warp-specialised tcgen05 core; interleaved-chains persistent mini-batch kernel
cd "${GRAFT_REPO_ROOT:-/root/repo}"
P=$PWD/resilient-consensus-based-marl_b200/rcmarl
mkdir -p gpurun_out
echo "== pytest -m gpu (default library)"
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -8
echo "== mini-batch chain: interleaved, exclusive shares, launch chain"
timeout 200 python tools/prof_mb.py 4096 3000 3 2>&1 | tail -2
RCMARL_MB_INTERLEAVE=0 timeout 200 python tools/prof_mb.py 4096 3000 3 2>&1 | tail -2
RCMARL_MB_PERSIST=0 timeout 200 python tools/prof_mb.py 4096 3000 3 2>&1 | tail -2
echo "== grad timing: default (tcgen05 WS) and the FFMA2 build"
timeout 200 python tools/ab_grad.py time 2>&1 | grep TIMING
RCMARL_LIB=$P/librcmarl_ffma.so timeout 200 python tools/ab_grad.py time 2>&1 | grep TIMING
echo "== bench (short)"
timeout 400 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-consensus > gpurun_out/bench_call9.json 2> gpurun_out/bench_call9.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/bench_call9.json"))
print("BENCH", d["value"], d["ms_per_step"], d["gpu_launches"], d["e2e"]["value"], json.dumps(d["roofline"]["regimes"]), json.dumps(d["breakdown_ms"]))
PY
tail -2 gpurun_out/bench_call9.err
echo "== ncu interleaved kernel"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:mb_persist_il -s 1 -c 1 -o gpurun_out/prof_mb_il python tools/prof_mb.py 4096 960 2 2>&1 | tail -2
