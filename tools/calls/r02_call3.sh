#!/bin/bash
# round 2, GPU call 3: full suite (team sizes, env options, persistent mini-batch kernel), grad A/B vs the pre-refactor build,
# ncu source capture of the tcgen05 hybrid (v6)
cd "${GRAFT_REPO_ROOT:-/root/repo}"
P=$PWD/resilient-consensus-based-marl_b200/rcmarl
mkdir -p gpurun_out
echo "== pytest -m gpu"
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -15
echo "== grad timing: current, pre-refactor (same box), current again"
timeout 200 python tools/ab_grad.py time 2>&1 | grep TIMING
RCMARL_LIB_LAX=1 RCMARL_LIB=$P/librcmarl_pre.so timeout 200 python tools/ab_grad.py time 2>&1 | grep TIMING
timeout 200 python tools/ab_grad.py time 2>&1 | grep TIMING
RCMARL_LIB=$P/librcmarl_v5.so timeout 200 python tools/ab_grad.py time 2>&1 | grep TIMING
echo "== ncu v6 (tcgen05 hybrid), source-level"
RCMARL_LIB=$P/librcmarl_v6.so timeout 600 ncu --set full --clock-control none --import-source on -k regex:grad_kernel_tc -s 2 -c 1 -o gpurun_out/prof_v6 python tools/prof_grad.py 4096000 8 3 2>&1 | tail -3
ls -la gpurun_out/*.ncu-rep
