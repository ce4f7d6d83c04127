#!/bin/bash
# round 2: team kernel launched per input kind (instruction-cache footprint), predicated LeakyReLU' multiply in the producers;
# DRAM traffic of a gradient launch in the bench's own setting (every job its own targets)
cd "${GRAFT_REPO_ROOT:-/root/repo}"
P=$PWD/resilient-consensus-based-marl_b200/rcmarl
echo "== pytest -m gpu"
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -4
echo "== timing"
timeout 200 python tools/prof_grad.py 12288000 8 5 2>&1 | tail -1
timeout 300 python tools/prof_mb.py 4096 3000 3 2>&1 | tail -1
echo "== bench (short)"
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_call23.json 2> gpurun_out/bench_call23.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/bench_call23.json"))
print("BENCH", d["value"], d["ms_per_step"], d["gpu_launches"], d["e2e"]["value"], d["breakdown_ms"])
PY
echo "== ncu: team kernel, gradient launch inside an update round"
NCU="ncu --set full --clock-control none --import-source on"
summ() { python tools/ncu_summary.py gpurun_out/r02_prof_$1.ncu-rep > gpurun_out/r02_ncu_$1.txt 2>&1; head -16 gpurun_out/r02_ncu_$1.txt | cut -c1-150; tail -3 gpurun_out/r02_ncu_$1.txt | cut -c1-200; rm -f gpurun_out/r02_prof_$1.ncu-rep; }
timeout 300 $NCU -k "regex:^team_kernel\$" -c 2 -f -o gpurun_out/r02_prof_team_kernel python tools/prof_round.py 2>&1 | tail -1; summ team_kernel
timeout 300 $NCU -k "regex:^grad_kernel_ws\$" -s 3 -c 1 -f -o gpurun_out/r02_prof_grad_ws_round python tools/prof_round.py 2>&1 | tail -1; summ grad_ws_round
