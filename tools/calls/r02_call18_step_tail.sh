#!/bin/bash
# round 2: step tail of the persistent mini-batch kernel -- padded park rows, back-off between cell polls (variants p0 / p100)
cd "${GRAFT_REPO_ROOT:-/root/repo}"
P=$PWD/resilient-consensus-based-marl_b200/rcmarl
timeout 600 python -m pytest tests/test_kernels_gpu.py tests/test_trainer_gpu.py -m gpu -q 2>&1 | tail -3
RCMARL_MB_TIMELINE=1 RCMARL_LIB=$P/librcmarl_tl.so timeout 300 python tools/prof_mb.py 4096 3000 2 2>&1 | tail -11
for v in "" _p0 _p100; do
  echo "== librcmarl$v.so"
  RCMARL_LIB=$P/librcmarl$v.so timeout 300 python tools/prof_mb.py 4096 3000 3 2>&1 | tail -1
done
timeout 200 python tools/prof_grad.py 12288000 8 5 2>&1 | tail -1
for v in _t1 _w1 _w2; do
  echo "== librcmarl$v.so"
  RCMARL_LIB=$P/librcmarl$v.so timeout 200 python tools/prof_grad.py 12288000 8 5 2>&1 | tail -1
  RCMARL_LIB=$P/librcmarl$v.so timeout 300 python tools/prof_mb.py 4096 3000 3 2>&1 | tail -1
done
echo "== t1: parity tests with the truncating split"
RCMARL_LIB=$P/librcmarl_t1.so timeout 600 python -m pytest tests/test_kernels_gpu.py tests/test_trainer_gpu.py tests/test_benchshape_parity_gpu.py -m gpu -q 2>&1 | tail -3
