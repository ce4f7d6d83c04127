#!/bin/bash
# One-call A/B on a GPU box: equal vs cost-balanced CTA shares of grad_kernel (RCMARL_BALANCED_GRID), then the GPU tests.
# Library build variants (csrc/Makefile `variants`) can be compared the same way with RCMARL_LIB=<path>.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
RCMARL_BALANCED_GRID=0 timeout 120 python tools/ab_grad.py dump gpurun_out/ab_uni.npz 2>&1 | tail -3
[ -f tools/ab_base_r01.npz ] && python tools/ab_grad.py cmp tools/ab_base_r01.npz gpurun_out/ab_uni.npz | tail -2
RCMARL_BALANCED_GRID=1 timeout 120 python tools/ab_grad.py dump gpurun_out/ab_bal.npz 2>&1 | tail -3
python tools/ab_grad.py cmp gpurun_out/ab_uni.npz gpurun_out/ab_bal.npz | tail -4
for m in 1 0; do
  echo "== bench RCMARL_BALANCED_GRID=$m"
  RCMARL_BALANCED_GRID=$m timeout 150 python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --no-consensus 2>&1 | tail -1 \
    | python -c "import sys, json; d = json.loads(sys.stdin.read()); print('BENCH', d['ms_per_step'], d['value'], d['gpu_launches'], d['roofline']['ms_per_launch'])"
done
echo "== pytest (defaults)"
timeout 300 python -m pytest tests -m gpu -q -x 2>&1 | tail -3
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
