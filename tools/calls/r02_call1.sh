#!/bin/bash
# round 2, GPU call 1: full GPU suite with the new parity tests, then the two never-run kernel variants (v5, v6)
cd "${GRAFT_REPO_ROOT:-/root/repo}"
P=$PWD/resilient-consensus-based-marl_b200/rcmarl
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv
echo "== pytest -m gpu (default library)"
timeout 900 python -m pytest tests -m gpu -q -x --durations=8 2>&1 | tail -25
echo "== default: dump + timing"
timeout 200 python tools/ab_grad.py dump gpurun_out/ab_base.npz 2>&1 | tail -3
for v in v6 v5; do
  echo "== library variant $v"
  RCMARL_LIB=$P/librcmarl_$v.so timeout 200 python tools/ab_grad.py dump gpurun_out/ab_$v.npz 2>&1 | tail -4
  python tools/ab_grad.py cmp gpurun_out/ab_base.npz gpurun_out/ab_$v.npz | tail -14
  echo "-- pytest kernels + trainer + benchshape with $v"
  RCMARL_LIB=$P/librcmarl_$v.so timeout 600 python -m pytest tests/test_kernels_gpu.py tests/test_trainer_gpu.py tests/test_benchshape_parity_gpu.py -m gpu -q 2>&1 | tail -12
done
echo "== bench (default, short)"
timeout 400 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_call1.json 2> gpurun_out/bench_call1.err; tail -c 3000 gpurun_out/bench_call1.json
echo "== smoke"
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
