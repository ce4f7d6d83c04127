"""Stand-alone driver for the mini-batch chain at the C2 shape (one rcmarl_minibatch_fit call = 940 SGD steps over
131 072-row mini-batches for the malicious agent's three chains), for ncu captures and A/B timing:
   python tools/prof_mb.py [n_envs=4096] [T=3000] [reps=3]          (RCMARL_MB_PERSIST=0: round-1 launch chain)
Prints CUDA-event timings per call and per step when run without a profiler."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "resilient-consensus-based-marl_b200"))
import bench                                           # noqa: E402
from rcmarl.trainer import Trainer                     # noqa: E402
from rcmarl import _lib as L                            # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
T = int(sys.argv[2]) if len(sys.argv) > 2 else 3000
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
cfg = bench.workload("C2", N)
tr = Trainer(seed=1, **cfg)
while tr.t_filled < T:
    tr.rollout_block(min(cfg["n_ep_fixed"], (T - tr.t_filled) // cfg["max_ep_len"]))
i = cfg["labels"].index("Malicious")
tr.tdt_local[0].normal_(); tr.tdt[i].normal_(); tr.neg_r_coop.normal_()
chains = [(tr.critic_local[i], L.IN_S, tr.tdt_local[0], 1, None), (tr.tr[i], L.IN_SA, tr.neg_r_coop, 1, tr.loss_t[i:i + 1]),
          (tr.critic[i], L.IN_S, tr.tdt[i], 1, tr.loss_c[i:i + 1])]
ts = []
for _ in range(reps):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); tr._minibatch_sgd(chains, tr.t_filled); b.record()
    torch.cuda.synchronize()
    ts.append(a.elapsed_time(b))
steps = tr.mb_epochs * ((tr.t_filled + tr.mb_times - 1) // tr.mb_times)
if os.environ.get("RCMARL_MB_TIMELINE") == "1":            # debug build (make variant_tl): per-step stage timeline of CTA 0
    import ctypes as C
    lib = L.lib()
    buf = (C.c_longlong * (64 * 16))()
    lib.rcmarl_debug_timeline_mb.restype = C.c_int
    st = lib.rcmarl_debug_timeline_mb(buf, 64 * 16)
    T = np.array(buf[:], np.int64).reshape(64, 16)[8:60]
    order = [0, 1, 2, 3, 4, 5, 6, 7, 8]                      # producer thread 0
    names = ["operand rebuild + barrier", "own tiles produced (sweep) + next step's first loads issued",
             "park (warp sums of the output-layer gradient)", "CTA barrier (consumers' last tile + their park)",
             "CTA sums -> level-1 cells", "gather (poll level-1, publish level-2)", "apply (poll level-2, SGD)", "closing barrier"]
    t = T[:, order]
    d = np.diff(t, axis=1)
    per = np.diff(T[:, 0])
    print("status", st, "step period (cycles): mean", per.mean(), "min", per.min(), "max", per.max())
    for k, n in enumerate(names):
        print(f"{n:62s} mean {d[:, k].mean():8.0f}  min {d[:, k].min():6d}  max {d[:, k].max():6d}")
    c = T[:, [1, 10, 11, 12]]                                # first consumer thread, relative to the producer's tick 1
    for n, a, b in (("consumer: last tile consumed, after the operands barrier", 0, 1), ("consumer: wait at the consumers' barrier", 1, 2),
                    ("consumer: park", 2, 3)):
        v = c[:, b] - c[:, a]
        print(f"{n:62s} mean {v.mean():8.0f}  min {v.min():6d}  max {v.max():6d}")
print(f"minibatch chain: n_envs={N} T={tr.t_filled} steps={steps} persistent={tr.mb_cells is not None} "
      f"ms_per_call={np.median(ts):.3f} us_per_step={1e3 * np.median(ts) / steps:.2f} (all: {[round(t, 2) for t in ts]})")
