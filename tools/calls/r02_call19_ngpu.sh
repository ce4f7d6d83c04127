#!/bin/bash
# round 2: weak scaling on all GPUs of the box (and 1 GPU on the same box)
cd "${GRAFT_REPO_ROOT:-/root/repo}"
N=$(nvidia-smi -L | wc -l); echo "GPUs: $N"
timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 3 --warmup 3 --no-cpu-baseline --no-consensus > gpurun_out/r02_bench_c2_${N}gpu.json 2> gpurun_out/r02_bench_c2_${N}gpu.err
tail -c 1200 gpurun_out/r02_bench_c2_${N}gpu.json; echo; tail -3 gpurun_out/r02_bench_c2_${N}gpu.err
timeout 300 python bench.py --gpus 1 --steps 3 --warmup 3 --no-cpu-baseline --no-consensus --no-e2e > gpurun_out/r02_bench_c2_1gpu_samebox${N}.json 2> /dev/null
tail -c 700 gpurun_out/r02_bench_c2_1gpu_samebox${N}.json; echo
