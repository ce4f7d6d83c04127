"""Full-size (BASELINE.json) checks through size-independent properties -- the oracle cannot run these sizes in seconds:
 * C5: clip-mean on 64 x 1M: permutation invariance of the neighbour rows, idempotence, H=0 == mean, bracketing;
 * C2: gradient sums over 4.1 M rows == sum of the sums over row shards (linearity of the reduction), bitwise
       determinism, and a local fit at full size strictly decreases its loss;
 * C2: the rollout block at 4096 envs is reproducible and independent of how the env batch is sharded."""
import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

from golden_util import pretrained                 # noqa: E402


def need_gpu():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")


@pytest.mark.parametrize("H", [0, 1, 4])
def test_c5_clip_mean_properties_at_full_size(H):
    need_gpu()
    from rcmarl import ops
    g = torch.Generator(device="cuda"); g.manual_seed(0)
    X = torch.randn(64, 1 << 20, device="cuda", generator=g)
    X[:, ::100] = torch.round(X[:, ::100])                      # ties on 1 % of the columns
    out = ops.clip_mean(X, H)
    perm = torch.cat([torch.zeros(1, dtype=torch.long, device="cuda"), 1 + torch.randperm(63, device="cuda", generator=g)])
    out_p = ops.clip_mean(X[perm].contiguous(), H)              # own row stays first, neighbours permuted
    assert float((out - out_p).abs().max()) < 5e-6 * float(X.abs().max())
    lo, hi = X.min(0).values, X.max(0).values
    assert bool(((out >= lo - 1e-6) & (out <= hi + 1e-6)).all())
    if H == 0:
        assert float((out - X.mean(0)).abs().max()) < 5e-6 * float(X.abs().max())
    else:
        # clipping can only pull the extremes towards the own value: |agg - own| <= |mean - own| is NOT guaranteed,
        # but the H=0 mean and the clipped mean coincide when all neighbours equal the own row
        same = X[:1].expand(64, -1).contiguous()
        assert float((ops.clip_mean(same, H) - X[0]).abs().max()) < 1e-6 * float(X.abs().max()) + 1e-6


def test_c2_gradient_sums_are_linear_in_the_row_set_and_deterministic():
    need_gpu()
    from rcmarl import ops, nets, _lib as L
    NA, N, T = 5, 4096, 1000
    B = N * T
    g = torch.Generator(device="cuda"); g.manual_seed(1)
    pos = torch.randint(0, 5, (B, NA, 2), device="cuda", generator=g).float()
    act = torch.randint(0, 5, (B, NA, 1), device="cuda", generator=g).float()
    sa = torch.cat([(pos - 2.0) / 1.41421354, act], -1).reshape(B, -1).contiguous()
    ns = ((torch.randint(0, 5, (B, NA, 2), device="cuda", generator=g).float() - 2.0) / 1.41421354).reshape(B, -1).contiguous()
    r = -torch.randint(0, 9, (B, NA), device="cuda", generator=g).float() / 5.0
    w, _, _ = pretrained()
    wc = torch.as_tensor(nets.pack(w[0][1])).cuda()
    wt = torch.as_tensor(nets.pack(w[0][2])).cuda()
    tgt = r[:, 0].contiguous()

    def sums(row_begin, n_rows):
        sc, st = torch.zeros(662, device="cuda"), torch.zeros(762, device="cuda")
        rows = ops.make_rows(sa, ns, r, NA, row_begin, n_rows)
        ops.grad(rows, [ops.grad_job(wc, tgt, sc, L.IN_S), ops.grad_job(wt, tgt, st, L.IN_SA)], L.LOSS_MSE)
        return sc, st
    full_c, full_t = sums(0, B)
    again_c, again_t = sums(0, B)
    assert torch.equal(full_c, again_c) and torch.equal(full_t, again_t)
    cut = 1234567                                                # ragged split, not a multiple of the tile size
    a_c, a_t = sums(0, cut)
    b_c, b_t = sums(cut, B - cut)
    for full, parts in ((full_c, a_c.double() + b_c.double()), (full_t, a_t.double() + b_t.double())):
        scale = float(full.abs().max())
        assert float((full.double() - parts).abs().max()) < 2e-5 * scale
    # one full-batch SGD step at full size lowers the loss (lr = 0.01, Keras MSE scaling)
    msg = torch.empty_like(wt)
    ops.sgd_apply([ops.sgd_job(msg, wt, full_t, 761, 0.01 * 2.0 / B)])
    st2 = torch.zeros(762, device="cuda")
    ops.grad(ops.make_rows(sa, ns, r, NA), [ops.grad_job(msg, tgt, st2, L.IN_SA)], L.LOSS_MSE)
    assert float(st2[-1]) < float(full_t[-1])


def test_c2_rollout_block_reproducible_and_shard_independent():
    need_gpu()
    from rcmarl.trainer import Trainer
    w, desired, labels = pretrained()
    kw = dict(labels=labels, in_nodes=[[0, 1, 2, 3], [1, 2, 3, 4], [2, 3, 4, 0], [3, 4, 0, 1], [4, 0, 1, 2]], weights=w,
              desired=desired, gamma=0.9, H=1, fast_lr=0.01, slow_lr=0.002, seed=77, buffer_size=50, n_ep_fixed=5)
    full = Trainer(n_envs=4096, **kw)
    e_full, r_full = full.rollout_block()
    again = Trainer(n_envs=4096, **kw)
    again.rollout_block()
    assert torch.equal(full.sa, again.sa) and torch.equal(full.r, again.r)
    half = Trainer(n_envs=2048, rank=1, world=1, **kw)           # rank only offsets the global env index here
    half.rank = 1
    half.rollout_block()
    T = 100
    assert torch.equal(full.sa.view(-1, 4096, 15)[:T, 2048:], half.sa.view(-1, 2048, 15)[:T])
    # rewards are non-positive multiples of 1/5, actions in range, est returns finite
    assert float(full.r.max()) <= 0 and np.isfinite(e_full).all() and (r_full <= 0).all()
