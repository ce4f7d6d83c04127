#!/bin/bash
# round 2, GPU call 6: warp-specialised kernel with input prefetch + sleeping waits (v7), three producer groups (v8)
cd "${GRAFT_REPO_ROOT:-/root/repo}"
P=$PWD/resilient-consensus-based-marl_b200/rcmarl
mkdir -p gpurun_out
echo "== pytest -m gpu"
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -6
echo "== grad timing: default, v7, v8, default"
timeout 200 python tools/ab_grad.py dump gpurun_out/ab_base.npz 2>&1 | grep -E "TIMING|rror"
for v in v7 v8; do
  RCMARL_LIB=$P/librcmarl_$v.so timeout 200 python tools/ab_grad.py dump gpurun_out/ab_$v.npz 2>&1 | grep -E "TIMING|rror|rap" | tail -4
  python tools/ab_grad.py cmp gpurun_out/ab_base.npz gpurun_out/ab_$v.npz | tail -3
done
echo "== v8: tests"
RCMARL_LIB=$P/librcmarl_v8.so timeout 600 python -m pytest tests/test_kernels_gpu.py tests/test_trainer_gpu.py tests/test_benchshape_parity_gpu.py -m gpu -q 2>&1 | tail -5
echo "== v7: tests"
RCMARL_LIB=$P/librcmarl_v7.so timeout 600 python -m pytest tests/test_kernels_gpu.py tests/test_trainer_gpu.py tests/test_benchshape_parity_gpu.py -m gpu -q 2>&1 | tail -5
echo "== ncu v7 + v8"
RCMARL_LIB=$P/librcmarl_v7.so timeout 600 ncu --set full --clock-control none --import-source on -k regex:grad_kernel_ws -s 2 -c 1 -o gpurun_out/prof_v7b python tools/prof_grad.py 4096000 8 3 2>&1 | tail -1
RCMARL_LIB=$P/librcmarl_v8.so timeout 600 ncu --set full --clock-control none --import-source on -k regex:grad_kernel_ws -s 2 -c 1 -o gpurun_out/prof_v8 python tools/prof_grad.py 4096000 8 3 2>&1 | tail -1
echo "== bench with v7 as the library (short) + clip-mean microbench with the default"
RCMARL_LIB=$P/librcmarl_v7.so timeout 300 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-consensus 2>/dev/null | python -c "import sys, json; d = json.loads(sys.stdin.read()); print('BENCH v7', d['value'], d['ms_per_step'], json.dumps(d['roofline']['regimes']), json.dumps(d['breakdown_ms']))"
RCMARL_LIB=$P/librcmarl_v8.so timeout 300 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-consensus 2>/dev/null | python -c "import sys, json; d = json.loads(sys.stdin.read()); print('BENCH v8', d['value'], d['ms_per_step'], json.dumps(d['roofline']['regimes']), json.dumps(d['breakdown_ms']))"
timeout 300 python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e 2>/dev/null | python -c "import sys, json; d = json.loads(sys.stdin.read()); print('BENCH default', d['value'], d['ms_per_step'], json.dumps({k: round(v['frac'], 3) for k, v in d['consensus_roofline'].items() if isinstance(v, dict)}))"
