#!/bin/bash
# round 2, evidence call: full GPU suite, benches (C2, reference arm, C3), ncu captures of every kernel at HEAD, launch list,
# compute-sanitizer, learning harness.  Everything lands in gpurun_out/ and is summarised into profiles/ afterwards.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
echo "== pytest -m gpu"
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -6
echo "== bench C2 (10 steps)"
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r02_bench_c2.json 2> gpurun_out/r02_bench_c2.err; tail -c 600 gpurun_out/r02_bench_c2.json; echo
echo "== bench reference arm"
timeout 600 python bench.py --impl reference --steps 3 --warmup 0 > gpurun_out/r02_bench_reference_arm.json 2> gpurun_out/r02_bench_reference_arm.err; tail -c 500 gpurun_out/r02_bench_reference_arm.json; echo
echo "== bench C3 (3 steps)"
timeout 900 python bench.py --workload C3 --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r02_bench_c3.json 2> gpurun_out/r02_bench_c3.err; tail -c 900 gpurun_out/r02_bench_c3.json; echo
echo "== ncu captures"
NCU="ncu --set full --clock-control none --import-source on"
summ() { python tools/ncu_summary.py gpurun_out/r02_prof_$1.ncu-rep > gpurun_out/r02_ncu_$1.txt 2>&1; head -12 gpurun_out/r02_ncu_$1.txt | cut -c1-150; [ "$2" = keep ] || rm -f gpurun_out/r02_prof_$1.ncu-rep; }
timeout 300 $NCU -k regex:grad_kernel_ws -s 2 -c 1 -o gpurun_out/r02_prof_grad_ws python tools/prof_grad.py 12288000 8 3 2>&1 | tail -1; summ grad_ws keep
timeout 400 $NCU -k regex:mb_persist_ws -s 1 -c 1 -o gpurun_out/r02_prof_mb_ws python tools/prof_mb.py 4096 960 2 2>&1 | tail -1; summ mb_ws keep
for k in team_kernel values_kernel rollout_kernel reduce_kernel consensus_hidden_kernel grad_kernel; do
  timeout 300 $NCU -k "regex:^$k\$" -c 1 -o gpurun_out/r02_prof_$k python tools/prof_round.py 2>&1 | tail -1; summ $k
done
for H in 1 2 4; do
  timeout 200 $NCU -k regex:clip_mean -s 2 -c 1 -o gpurun_out/r02_prof_clip_h$H python tools/prof_clip_mean.py $H 2>&1 | tail -1; summ clip_h$H
done
echo "== launch list of one bench step"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/r02_launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --no-consensus > /dev/null 2>&1; wc -l gpurun_out/r02_launches.csv
echo "== compute-sanitizer (memcheck, racecheck) over smoke()"
timeout 400 compute-sanitizer --tool memcheck python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -4 | tee gpurun_out/r02_sanitizer_memcheck.log
timeout 400 compute-sanitizer --tool racecheck python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -4 | tee gpurun_out/r02_sanitizer_racecheck.log
echo "== learning harness"
timeout 800 python tools/learning_harness.py 30 256 2>&1 | tail -30
ls -la gpurun_out | tail -30
