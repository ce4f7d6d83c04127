#!/bin/bash
# round 2, final evidence at HEAD: GPU suite, C2 bench (10 steps), ncu captures of the two tensor-core kernels, launch list,
# producer / step timelines, same-box timing of the default and the FFMA2 build
cd "${GRAFT_REPO_ROOT:-/root/repo}"
P=$PWD/resilient-consensus-based-marl_b200/rcmarl
mkdir -p gpurun_out
echo "== pytest -m gpu"
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -4
echo "== smoke"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
echo "== bench C2 (10 steps)"
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r02_bench_c2.json 2> gpurun_out/r02_bench_c2.err; tail -c 400 gpurun_out/r02_bench_c2.json; echo
echo "== timing: default, FFMA2 build"
for v in "" _ffma; do
  RCMARL_LIB=$P/librcmarl$v.so timeout 200 python tools/prof_grad.py 12288000 8 5 2>&1 | tail -1
  RCMARL_LIB=$P/librcmarl$v.so timeout 300 python tools/prof_mb.py 4096 3000 3 2>&1 | tail -1
done
echo "== timelines (debug build)"
RCMARL_LIB=$P/librcmarl_tl.so timeout 200 python tools/ws_timeline.py 2>&1 | tail -12
RCMARL_MB_TIMELINE=1 RCMARL_LIB=$P/librcmarl_tl.so timeout 300 python tools/prof_mb.py 4096 3000 2 2>&1 | tail -14
echo "== ncu captures"
NCU="ncu --set full --clock-control none --import-source on"
summ() { python tools/ncu_summary.py gpurun_out/r02_prof_$1.ncu-rep > gpurun_out/r02_ncu_$1.txt 2>&1; head -14 gpurun_out/r02_ncu_$1.txt | cut -c1-150; }
timeout 300 $NCU -k regex:grad_kernel_ws -s 2 -c 1 -f -o gpurun_out/r02_prof_grad_ws python tools/prof_grad.py 12288000 8 3 2>&1 | tail -1; summ grad_ws
timeout 400 $NCU -k regex:mb_persist_ws -s 1 -c 1 -f -o gpurun_out/r02_prof_mb_ws python tools/prof_mb.py 4096 960 2 2>&1 | tail -1; summ mb_ws
echo "== launch list of the bench command"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/r02_launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --no-consensus > /dev/null 2>&1; wc -l gpurun_out/r02_launches.csv
