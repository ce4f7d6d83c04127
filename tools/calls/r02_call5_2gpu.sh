#!/bin/bash
# round 2, GPU call 5 (2 GPUs): data-parallel equivalence tests (LL exchange, 5- and 16-agent teams) and the 2-GPU bench
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
nvidia-smi -L
echo "== pytest test_dp_gpu"
timeout 600 python -m pytest tests/test_dp_gpu.py -m gpu -q 2>&1 | tail -8
echo "== bench 2 GPUs"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 --no-cpu-baseline --no-consensus > gpurun_out/bench_2gpu_r02.json 2> gpurun_out/bench_2gpu_r02.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/bench_2gpu_r02.json"))
print("BENCH2", d["value"], d["ms_per_step"], d["replicas_identical"], json.dumps(d["breakdown_ms"]), d["roofline"]["regimes"]["mini_batch"]["us_per_step"])
PY
tail -3 gpurun_out/bench_2gpu_r02.err
echo "== bench 1 GPU (same box)"
timeout 400 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-consensus --no-e2e > gpurun_out/bench_1gpu_r02.json 2> gpurun_out/bench_1gpu_r02.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/bench_1gpu_r02.json"))
print("BENCH1", d["value"], d["ms_per_step"], json.dumps(d["breakdown_ms"]), d["roofline"]["regimes"]["mini_batch"]["us_per_step"])
PY
