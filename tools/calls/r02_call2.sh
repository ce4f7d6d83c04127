#!/bin/bash
# round 2, GPU call 2: persistent mini-batch kernel (tests, A/B timing, bench), ncu source capture of the tcgen05 hybrid
cd "${GRAFT_REPO_ROOT:-/root/repo}"
P=$PWD/resilient-consensus-based-marl_b200/rcmarl
mkdir -p gpurun_out
echo "== pytest -m gpu"
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -15
echo "== mini-batch chain A/B"
timeout 200 python tools/prof_mb.py 4096 3000 3 2>&1 | tail -2
RCMARL_MB_PERSIST=0 timeout 200 python tools/prof_mb.py 4096 3000 3 2>&1 | tail -2
echo "== bench short (persistent)"
timeout 400 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-consensus > gpurun_out/bench_call2.json 2> gpurun_out/bench_call2.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/bench_call2.json"))
print("BENCH", d["value"], d["ms_per_step"], d["gpu_launches"], d["e2e"]["value"], json.dumps(d["roofline"]["regimes"]), json.dumps(d["breakdown_ms"]))
PY
echo "== ncu v6 (tcgen05 hybrid), source-level"
RCMARL_LIB=$P/librcmarl_v6.so timeout 600 ncu --set full --clock-control none --import-source on -k regex:grad_kernel_tc -s 2 -c 1 -o gpurun_out/prof_v6 python tools/prof_grad.py 4096000 8 3 2>&1 | tail -3
echo "== ncu persistent kernel"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:mb_persist -s 1 -c 1 -o gpurun_out/prof_mb python tools/prof_mb.py 4096 960 2 2>&1 | tail -3
ls -la gpurun_out/*.ncu-rep
