#!/bin/bash
# round 2: two-rank tests and a short two-rank bench at the final HEAD (team kernel split, prefetch, poll back-off)
cd "${GRAFT_REPO_ROOT:-/root/repo}"
N=$(nvidia-smi -L | wc -l); echo "GPUs: $N"
timeout 600 python -m pytest tests/test_dp_gpu.py -m gpu -q 2>&1 | tail -3
timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus $N --steps 3 --warmup 3 --no-cpu-baseline --no-consensus > gpurun_out/r02_bench_c2_${N}gpu_final.json 2> gpurun_out/r02_bench_c2_${N}gpu_final.err
python - <<PY
import json
d = json.load(open("gpurun_out/r02_bench_c2_${N}gpu_final.json"))
print("BENCH", d["n_gpus"], d["value"], d["ms_per_step"], d["replicas_identical"], d["e2e"]["value"], d["breakdown_ms"])
PY
tail -2 gpurun_out/r02_bench_c2_${N}gpu_final.err
