#!/bin/bash
# round 2: sweep of the producers' MMA-wait back-off (csrc/Makefile variants s0..s5), full-batch and mini-batch regimes
cd "${GRAFT_REPO_ROOT:-/root/repo}"
P=resilient-consensus-based-marl_b200/rcmarl
echo "== timeline with immediate polling (true MMA round-trip latencies)"
RCMARL_LIB=$PWD/$P/librcmarl_tl0.so timeout 200 python tools/ws_timeline.py 2>&1 | tail -14
for v in "" _s0 _s1 _s2 _s3 _s4 _s5; do
  echo "== librcmarl$v.so"
  RCMARL_LIB=$PWD/$P/librcmarl$v.so timeout 200 python tools/prof_grad.py 4096000 8 5 2>&1 | tail -2
  RCMARL_LIB=$PWD/$P/librcmarl$v.so timeout 300 python tools/prof_mb.py 4096 960 2 2>&1 | tail -2
done
