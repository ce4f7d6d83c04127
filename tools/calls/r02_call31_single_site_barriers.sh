#!/bin/bash
# round 2: CTA-wide barriers behind the role branch (one program location) -- GPU suite, timings, synccheck
cd "${GRAFT_REPO_ROOT:-/root/repo}"
timeout 600 python -m pytest tests -m gpu -q -x 2>&1 | tail -3
timeout 150 python tools/prof_grad.py 12288000 8 5 2>&1 | tail -1
timeout 200 python tools/prof_mb.py 4096 3000 3 2>&1 | tail -1
timeout 200 compute-sanitizer --tool synccheck python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
