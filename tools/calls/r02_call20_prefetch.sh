#!/bin/bash
# round 2: next step's first rows fetched ahead of the step tail; finer step timeline (producer park, first consumer thread)
cd "${GRAFT_REPO_ROOT:-/root/repo}"
P=$PWD/resilient-consensus-based-marl_b200/rcmarl
timeout 600 python -m pytest tests/test_kernels_gpu.py tests/test_trainer_gpu.py -m gpu -q 2>&1 | tail -3
RCMARL_MB_TIMELINE=1 RCMARL_LIB=$P/librcmarl_tl.so timeout 300 python tools/prof_mb.py 4096 3000 2 2>&1 | tail -14
for v in "" _prev ""; do
  echo "== librcmarl$v.so"
  RCMARL_LIB=$P/librcmarl$v.so timeout 300 python tools/prof_mb.py 4096 3000 3 2>&1 | tail -1
done
