#!/bin/bash
# round 2, GPU call 4: the warp-specialised tcgen05 kernel (variant v7): numerics vs the default, timing, tests, ncu
cd "${GRAFT_REPO_ROOT:-/root/repo}"
P=$PWD/resilient-consensus-based-marl_b200/rcmarl
mkdir -p gpurun_out
echo "== v7: dump / compare / timing"
timeout 200 python tools/ab_grad.py dump gpurun_out/ab_base.npz 2>&1 | grep -E "dumped|TIMING|rror"
RCMARL_LIB=$P/librcmarl_v7.so timeout 200 python tools/ab_grad.py dump gpurun_out/ab_v7.npz 2>&1 | grep -E "dumped|TIMING|rror|rap" | tail -5
python tools/ab_grad.py cmp gpurun_out/ab_base.npz gpurun_out/ab_v7.npz | tail -10
echo "== v7: tests"
RCMARL_LIB=$P/librcmarl_v7.so timeout 600 python -m pytest tests/test_kernels_gpu.py tests/test_trainer_gpu.py tests/test_benchshape_parity_gpu.py -m gpu -q -x 2>&1 | tail -12
echo "== v7: ncu"
RCMARL_LIB=$P/librcmarl_v7.so timeout 600 ncu --set full --clock-control none --import-source on -k regex:grad_kernel_ws -s 2 -c 1 -o gpurun_out/prof_v7 python tools/prof_grad.py 4096000 8 3 2>&1 | tail -3
RCMARL_LIB=$P/librcmarl_v7.so timeout 100 python tools/prof_grad.py 12288000 8 5 2>&1 | tail -1
ls -la gpurun_out/*.ncu-rep
