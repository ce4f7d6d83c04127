#!/bin/bash
# round 2: per-step stage timeline of the persistent mini-batch kernel (debug build), default timing
cd "${GRAFT_REPO_ROOT:-/root/repo}"
P=$PWD/resilient-consensus-based-marl_b200/rcmarl
RCMARL_MB_TIMELINE=1 RCMARL_LIB=$P/librcmarl_tl.so timeout 300 python tools/prof_mb.py 4096 3000 2 2>&1 | tail -11
timeout 300 python tools/prof_mb.py 4096 3000 3 2>&1 | tail -1
timeout 200 python tools/prof_grad.py 12288000 8 5 2>&1 | tail -1
