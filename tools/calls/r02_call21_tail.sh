#!/bin/bash
# round 2: pipelined park reduction; fewer / slower pollers in the apply stage (variants a128, a128s)
cd "${GRAFT_REPO_ROOT:-/root/repo}"
P=$PWD/resilient-consensus-based-marl_b200/rcmarl
timeout 600 python -m pytest tests/test_kernels_gpu.py tests/test_trainer_gpu.py -m gpu -q 2>&1 | tail -3
RCMARL_MB_TIMELINE=1 RCMARL_LIB=$P/librcmarl_tl.so timeout 300 python tools/prof_mb.py 4096 3000 2 2>&1 | tail -14
for v in "" _prev _a128 _a128s; do
  echo "== librcmarl$v.so"
  RCMARL_LIB=$P/librcmarl$v.so timeout 300 python tools/prof_mb.py 4096 3000 3 2>&1 | tail -1
done
RCMARL_LIB=$P/librcmarl_a128.so timeout 600 python -m pytest tests/test_kernels_gpu.py tests/test_trainer_gpu.py -m gpu -q 2>&1 | tail -3
timeout 200 python tools/prof_grad.py 12288000 8 5 2>&1 | tail -1
