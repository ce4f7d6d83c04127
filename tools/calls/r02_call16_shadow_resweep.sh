#!/bin/bash
# round 2: deferral-mask re-sweep after the elect.sync issue change (csrc/Makefile variants q0..q7), both regimes
cd "${GRAFT_REPO_ROOT:-/root/repo}"
P=$PWD/resilient-consensus-based-marl_b200/rcmarl
for v in "" _q0 _q2 _q4 _q6 _q7; do
  echo "== librcmarl$v.so"
  RCMARL_LIB=$P/librcmarl$v.so timeout 200 python tools/prof_grad.py 12288000 8 5 2>&1 | tail -1
  RCMARL_LIB=$P/librcmarl$v.so timeout 300 python tools/prof_mb.py 4096 3000 3 2>&1 | tail -1
done
