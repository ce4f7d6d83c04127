#!/bin/bash
# round 2: which producer bookkeeping hides behind the MMA batches (csrc/Makefile variants m0..m6), both regimes
cd "${GRAFT_REPO_ROOT:-/root/repo}"
P=$PWD/resilient-consensus-based-marl_b200/rcmarl
for v in "" _m0 _m2 _m3 _m4 _m6 ""; do
  echo "== librcmarl$v.so"
  RCMARL_LIB=$P/librcmarl$v.so timeout 200 python tools/prof_grad.py 12288000 8 5 2>&1 | tail -1
  RCMARL_LIB=$P/librcmarl$v.so timeout 300 python tools/prof_mb.py 4096 3000 3 2>&1 | tail -1
done
