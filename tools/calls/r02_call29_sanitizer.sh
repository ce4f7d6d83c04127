#!/bin/bash
# round 2: compute-sanitizer memcheck and synccheck over smoke() at HEAD
cd "${GRAFT_REPO_ROOT:-/root/repo}"
timeout 400 compute-sanitizer --tool memcheck python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | tee gpurun_out/r02_sanitizer_memcheck.log
timeout 400 compute-sanitizer --tool synccheck python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -6 | tee gpurun_out/r02_sanitizer_synccheck.log
timeout 400 compute-sanitizer --tool initcheck python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -6 | tee gpurun_out/r02_sanitizer_initcheck.log
