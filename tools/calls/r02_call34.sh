#!/bin/bash
# round 2, GPU calls 3+4 merged: full suite, grad A/B vs the pre-refactor build, v7 (warp-specialised tcgen05 kernel)
cd "${GRAFT_REPO_ROOT:-/root/repo}"
P=$PWD/resilient-consensus-based-marl_b200/rcmarl
mkdir -p gpurun_out
echo "== pytest -m gpu"
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -15
echo "== grad timing: current, pre-refactor (same box), current again, v5"
timeout 200 python tools/ab_grad.py time 2>&1 | grep TIMING
RCMARL_LIB_LAX=1 RCMARL_LIB=$P/librcmarl_pre.so timeout 200 python tools/ab_grad.py time 2>&1 | grep TIMING
timeout 200 python tools/ab_grad.py time 2>&1 | grep TIMING
echo "== v7: dump / compare / timing"
timeout 200 python tools/ab_grad.py dump gpurun_out/ab_base.npz 2>&1 | grep -E "dumped|rror"
RCMARL_LIB=$P/librcmarl_v7.so timeout 200 python tools/ab_grad.py dump gpurun_out/ab_v7.npz 2>&1 | grep -E "dumped|TIMING|rror|rap" | tail -5
python tools/ab_grad.py cmp gpurun_out/ab_base.npz gpurun_out/ab_v7.npz | tail -10
echo "== v7: tests"
RCMARL_LIB=$P/librcmarl_v7.so timeout 600 python -m pytest tests/test_kernels_gpu.py tests/test_trainer_gpu.py tests/test_benchshape_parity_gpu.py -m gpu -q -x 2>&1 | tail -12
echo "== v7: ncu"
RCMARL_LIB=$P/librcmarl_v7.so timeout 600 ncu --set full --clock-control none --import-source on -k regex:grad_kernel_ws -s 2 -c 1 -o gpurun_out/prof_v7 python tools/prof_grad.py 4096000 8 3 2>&1 | tail -2
echo "== clip-mean microbench"
timeout 300 python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench_call34.json 2> gpurun_out/bench_call34.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/bench_call34.json"))
print("BENCH", d["value"], d["ms_per_step"], json.dumps({k: round(v["frac"], 3) for k, v in d["consensus_roofline"].items() if isinstance(v, dict)}))
PY
ls -la gpurun_out/*.ncu-rep
