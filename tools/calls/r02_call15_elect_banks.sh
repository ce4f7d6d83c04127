#!/bin/bash
# round 2: conflict-free buffer rows + elect.sync MMA issue -- tests, timeline, both regimes, bench; racecheck report in full;
# ncu of the 16-agent gradient kernel at the C3 shape
cd "${GRAFT_REPO_ROOT:-/root/repo}"
P=$PWD/resilient-consensus-based-marl_b200/rcmarl
echo "== pytest -m gpu"
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -4
echo "== timeline"
RCMARL_LIB=$P/librcmarl_tl.so timeout 200 python tools/ws_timeline.py 2>&1 | tail -12
echo "== timing"
timeout 200 python tools/prof_grad.py 12288000 8 5 2>&1 | tail -1
timeout 300 python tools/prof_mb.py 4096 3000 3 2>&1 | tail -1
echo "== bench (short)"
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_call15.json 2> gpurun_out/bench_call15.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/bench_call15.json"))
print("BENCH", d["value"], d["ms_per_step"], d["gpu_launches"], d["e2e"]["value"], json.dumps(d["roofline"]["regimes"])[:700])
PY
echo "== racecheck, full report"
timeout 400 compute-sanitizer --tool racecheck --racecheck-report all python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_sanitizer_racecheck_full.log 2>&1; grep -c "" gpurun_out/r02_sanitizer_racecheck_full.log; grep -E "Race reported|hazard" gpurun_out/r02_sanitizer_racecheck_full.log | sed 's/+0x[0-9a-f]*//' | sort | uniq -c | sort -rn | head -30
RCMARL_LIB=$P/librcmarl_ffma.so timeout 400 compute-sanitizer --tool racecheck python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
echo "== ncu: 16-agent gradient kernel at the C3 shape"
timeout 600 ncu --set full --clock-control none --import-source on -k "regex:^grad_kernel\$" -c 1 -o gpurun_out/r02_prof_grad16 python tools/prof_round.py C3 4096 2>&1 | tail -1
python tools/ncu_summary.py gpurun_out/r02_prof_grad16.ncu-rep > gpurun_out/r02_ncu_grad16.txt 2>&1; head -30 gpurun_out/r02_ncu_grad16.txt | cut -c1-150; tail -16 gpurun_out/r02_ncu_grad16.txt | cut -c1-200
