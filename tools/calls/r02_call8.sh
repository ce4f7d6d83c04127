#!/bin/bash
# round 2, GPU call 8: v7 = warp-specialised tcgen05 core, consumers share every tile, + persistent mini-batch kernel on it
cd "${GRAFT_REPO_ROOT:-/root/repo}"
P=$PWD/resilient-consensus-based-marl_b200/rcmarl
mkdir -p gpurun_out
echo "== grad timing: default, v7"
timeout 200 python tools/ab_grad.py dump gpurun_out/ab_base.npz 2>&1 | grep -E "TIMING|rror"
RCMARL_LIB=$P/librcmarl_v7.so timeout 200 python tools/ab_grad.py dump gpurun_out/ab_v7.npz 2>&1 | grep -E "TIMING|rror|rap" | tail -4
python tools/ab_grad.py cmp gpurun_out/ab_base.npz gpurun_out/ab_v7.npz | tail -3
echo "== v7: tests"
RCMARL_LIB=$P/librcmarl_v7.so timeout 600 python -m pytest tests/test_kernels_gpu.py tests/test_trainer_gpu.py tests/test_benchshape_parity_gpu.py tests/test_api_gpu.py -m gpu -q 2>&1 | tail -8
echo "== mini-batch chain with v7 (persistent WS kernel), then launch chain"
RCMARL_LIB=$P/librcmarl_v7.so timeout 200 python tools/prof_mb.py 4096 3000 3 2>&1 | tail -2
RCMARL_LIB=$P/librcmarl_v7.so RCMARL_MB_PERSIST=0 timeout 200 python tools/prof_mb.py 4096 3000 3 2>&1 | tail -2
echo "== ncu v7"
RCMARL_LIB=$P/librcmarl_v7.so timeout 600 ncu --set full --clock-control none --import-source on -k regex:grad_kernel_ws -s 2 -c 1 -o gpurun_out/prof_v7f python tools/prof_grad.py 4096000 8 3 2>&1 | tail -1
echo "== bench with v7 (short)"
RCMARL_LIB=$P/librcmarl_v7.so timeout 300 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-consensus 2>/dev/null | python -c "import sys, json; d = json.loads(sys.stdin.read()); print('BENCH v7', d['value'], d['ms_per_step'], json.dumps(d['roofline']['regimes']), json.dumps(d['breakdown_ms']))"
