"""A/B check of librcmarl build variants (csrc/Makefile `variants`): dump the gradient sums of fixed seeded inputs with the
library selected by RCMARL_LIB, then compare two dumps bit for bit.

    RCMARL_LIB=.../librcmarl.so    python tools/ab_grad.py dump gpurun_out/ab_base.npz
    RCMARL_LIB=.../librcmarl_v2.so python tools/ab_grad.py dump gpurun_out/ab_v2.npz
    python tools/ab_grad.py cmp gpurun_out/ab_base.npz gpurun_out/ab_v2.npz

The variants only change instruction selection and where operands come from, never the order of the floating-point
operations, so the dumps must be identical."""
import os
import sys

import numpy as np


def dump(path):
    import torch
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, os.path.join(root, "resilient-consensus-based-marl_b200"))
    from rcmarl import ops, nets, _lib as L
    dev = torch.device("cuda:0")
    out = {}
    for na, n_rows in ((5, (1 << 20) + 77), (16, (1 << 18) + 13)):
        g = torch.Generator(device="cuda"); g.manual_seed(na)
        sa = torch.randn(n_rows, 3 * na, device=dev, generator=g)
        sa[:, 2::3] = torch.randint(0, 5, (n_rows, na), device=dev, generator=g).float()     # action slots
        ns = torch.randn(n_rows, 2 * na, device=dev, generator=g)
        r = torch.randn(n_rows, na, device=dev, generator=g)
        tgt = torch.randn(n_rows, device=dev, generator=g)
        rows = ops.make_rows(sa, ns, r, na)
        rs = np.random.RandomState(na)
        kinds = [L.IN_SA, L.IN_S, L.IN_NS, L.IN_SA, L.IN_S, L.IN_NS, L.IN_S, L.IN_SA]
        jobs, sums = [], []
        for kind in kinds:
            din = 3 * na if kind == L.IN_SA else 2 * na
            w = torch.as_tensor(nets.pack(nets.glorot_uniform(din, 1, rs))).to(dev)
            s = torch.zeros(L.param_count(din, 1) + 1, device=dev)
            jobs.append(ops.grad_job(w, tgt, s, kind))
            sums.append(s)
        ops.grad(rows, jobs, L.LOSS_MSE)
        jobs_ce, sums_ce = [], []
        for a in range(4):
            w = torch.as_tensor(nets.pack(nets.glorot_uniform(2 * na, 5, rs))).to(dev)
            s = torch.zeros(L.param_count(2 * na, 5) + 1, device=dev)
            jobs_ce.append(ops.grad_job(w, tgt, s, L.IN_S, action_agent=a))
            sums_ce.append(s)
        ops.grad(rows, jobs_ce, L.LOSS_CE)
        torch.cuda.synchronize()
        for i, s in enumerate(sums):
            out[f"na{na}_mse{i}"] = s.cpu().numpy()
        for i, s in enumerate(sums_ce):
            out[f"na{na}_ce{i}"] = s.cpu().numpy()
    np.savez(path, **out)
    print("dumped", len(out), "arrays to", path, "lib:", os.environ.get("RCMARL_LIB", "(default)"))


def timing():
    """CUDA-event timings of rcmarl_grad (+ its reduce) at the C2 full-batch and mini-batch shapes."""
    import torch
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, os.path.join(root, "resilient-consensus-based-marl_b200"))
    from rcmarl import ops, nets, _lib as L
    dev = torch.device("cuda:0")
    na = 5
    for n_rows, n_jobs, reps in ((12288000, 8, 9), (131072, 3, 300)):
        g = torch.Generator(device="cuda"); g.manual_seed(0)
        sa = torch.randn(n_rows, 3 * na, device=dev, generator=g)
        ns = torch.randn(n_rows, 2 * na, device=dev, generator=g)
        r = torch.randn(n_rows, na, device=dev, generator=g)
        tgt = torch.randn(n_rows, device=dev, generator=g)
        rows = ops.make_rows(sa, ns, r, na)
        rs = np.random.RandomState(0)
        jobs, keep = [], []
        for j in range(n_jobs):
            kind = L.IN_SA if j % 2 == 0 else L.IN_S
            din = 15 if kind == L.IN_SA else 10
            w = torch.as_tensor(nets.pack(nets.glorot_uniform(din, 1, rs))).to(dev)
            s = torch.zeros(L.param_count(din, 1) + 1, device=dev)
            jobs.append(ops.grad_job(w, tgt, s, kind)); keep.append((w, s))
        jobs = (L.GradJob * n_jobs)(*jobs)
        for _ in range(3):
            ops.grad(rows, jobs, L.LOSS_MSE)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(reps):
            ops.grad(rows, jobs, L.LOSS_MSE)
        b.record()
        torch.cuda.synchronize()
        ms = a.elapsed_time(b) / reps
        print(f"TIMING lib={os.path.basename(os.environ.get('RCMARL_LIB', 'librcmarl.so'))} rows={n_rows} jobs={n_jobs} "
              f"ms_per_launch={ms:.4f}", flush=True)
        del sa, ns, r, tgt


def cmp(a, b):
    A, B = np.load(a), np.load(b)
    bad = 0
    for k in A.files:
        same = np.array_equal(A[k].view(np.uint32), B[k].view(np.uint32))
        if not same:
            bad += 1
            print(f"{k}: DIFFERENT  max abs diff {np.abs(A[k] - B[k]).max():.3e}  (max |a| {np.abs(A[k]).max():.3e})")
    print(f"{a} vs {b}: {len(A.files) - bad}/{len(A.files)} arrays bit-identical")
    return bad


if __name__ == "__main__":
    if sys.argv[1] == "time":
        timing()
    elif sys.argv[1] == "dump":
        dump(sys.argv[2])
        timing()
    else:
        sys.exit(1 if cmp(sys.argv[2], sys.argv[3]) else 0)
