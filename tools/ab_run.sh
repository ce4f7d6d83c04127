#!/bin/bash
# One-call A/B of the librcmarl build variants on a GPU box (csrc/Makefile `variants`, tools/ab_grad.py):
# bit-for-bit comparison of the gradient sums, kernel timings, a short C2 bench per library, then the GPU tests.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
P=$PWD/resilient-consensus-based-marl_b200/rcmarl
mkdir -p gpurun_out
VARIANTS="${VARIANTS:-v0 v4}"
RCMARL_LIB=$P/librcmarl.so timeout 120 python tools/ab_grad.py dump gpurun_out/ab_base.npz 2>&1 | tail -3
for v in $VARIANTS; do
  RCMARL_LIB=$P/librcmarl_$v.so timeout 120 python tools/ab_grad.py dump gpurun_out/ab_$v.npz 2>&1 | tail -3
  python tools/ab_grad.py cmp gpurun_out/ab_base.npz gpurun_out/ab_$v.npz
done
for v in base $VARIANTS; do
  lib=$P/librcmarl.so; [ $v != base ] && lib=$P/librcmarl_$v.so
  echo "== bench $v"
  RCMARL_LIB=$lib timeout 150 python bench.py --steps 2 --warmup 2 --no-cpu-baseline --no-e2e --no-consensus 2>&1 | tail -1 \
    | python -c "import sys, json; d = json.loads(sys.stdin.read()); print('BENCH', d['ms_per_step'], d['value'], d['gpu_launches'])"
done
echo "== pytest (default library)"
timeout 300 python -m pytest tests -m gpu -q -x 2>&1 | tail -3
for v in $TEST_VARIANTS; do
  echo "== pytest kernels + trainer ($v)"
  RCMARL_LIB=$P/librcmarl_$v.so timeout 200 python -m pytest tests/test_kernels_gpu.py tests/test_trainer_gpu.py -m gpu -q -x 2>&1 | tail -3
done
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
