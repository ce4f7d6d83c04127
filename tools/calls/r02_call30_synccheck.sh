#!/bin/bash
# round 2: synccheck report in full (first records + histogram)
cd "${GRAFT_REPO_ROOT:-/root/repo}"
timeout 400 compute-sanitizer --tool synccheck --print-limit 40 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_sanitizer_synccheck_full.log 2>&1
grep -c "" gpurun_out/r02_sanitizer_synccheck_full.log
grep -E "=========( Error:| Barrier|     at )" gpurun_out/r02_sanitizer_synccheck_full.log | sed -E 's/\+0x[0-9a-f]+//; s/thread \([0-9]+,0,0\)/thread (T)/; s/block \([0-9]+,0,0\)/block (B)/' | sort | uniq -c | sort -rn | head -20
head -30 gpurun_out/r02_sanitizer_synccheck_full.log | cut -c1-220
RCMARL_LIB=$PWD/resilient-consensus-based-marl_b200/rcmarl/librcmarl_ffma.so timeout 400 compute-sanitizer --tool synccheck python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
