"""One small invocation of the whole hot path on cuda:0, checked against the CPU oracle
(the only place outside tests/ and bench.py's CPU arm where oracle/ is imported: __graft_entry__.smoke())."""
import os
import sys

import numpy as np
import torch


def run(verbose=False):
    root = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    if root not in sys.path:
        sys.path.insert(0, root)
    from oracle import rpbcac_oracle as O                 # checker only
    from rcmarl.trainer import Trainer
    z = np.load(os.path.join(root, "tests", "golden", "kat_est_returns.npz"))
    tag = "malicious_H1_s100"
    w = [[[z[f"{tag}/agent{i}/n{n}_k{k}"] for k in range(6)] for n in range(4 if i == 4 else 3)] for i in range(5)]
    desired, labels = z[f"{tag}/desired"], [str(x) for x in z[f"{tag}/labels"]]
    in_nodes = [[0, 1, 2, 3], [1, 2, 3, 4], [2, 3, 4, 0], [3, 4, 0, 1], [4, 0, 1, 2]]
    N, n_ep, Lq, gamma = 8, 4, 10, 0.9
    rs = np.random.RandomState(0)
    init = rs.randint(0, 5, size=(n_ep, N, 5, 2)).astype(np.int32)
    U = rs.rand(n_ep, Lq, N, 5, 3).astype(np.float32)
    prs = np.random.RandomState(1)
    used = []

    def rec(T):
        p = prs.permutation(T)
        used.append(p)
        return p
    agents = []
    for i, l in enumerate(labels):
        if l == "Malicious":
            agents.append(O.MaliciousOracleAgent(w[i][0], w[i][1], w[i][2], 0.002, 0.01, gamma, critic_local_w=w[i][3], dtype=np.float64))
        else:
            agents.append(O.RPBCACOracleAgent(w[i][0], w[i][1], w[i][2], 0.002, 0.01, gamma, H=1, dtype=np.float64))
    env = O.GridWorldOracle(5, 5, 5, desired, n_envs=N)
    S, NS, A, R, est, ret = O.rollout_block(env, agents, labels, n_episodes=n_ep, max_ep_len=Lq, gamma=gamma,
                                            init_states=init, uniforms=U)
    O.update_round(agents, labels, in_nodes, S, NS, A, R, n_envs=N, n_epochs=1, n_actor_steps=n_ep * Lq,
                   common_reward=False, perm_source=rec)
    it = iter(list(used))
    tr = Trainer(labels=labels, in_nodes=in_nodes, weights=w, desired=desired, n_envs=N, gamma=gamma, H=1, fast_lr=0.01,
                 slow_lr=0.002, max_ep_len=Lq, n_ep_fixed=n_ep, n_epochs=1, buffer_size=1000, perm_source=lambda T: next(it))
    dev = tr.dev
    g_est, g_ret = tr.rollout_block(n_ep, uniforms=torch.as_tensor(U).to(dev), init_state=torch.as_tensor(init).to(dev))
    same = (tr.sa[:n_ep * Lq * N].cpu().numpy().reshape(-1, 5, 3)[:, :, 2] == A[:, :, 0]).mean()
    assert same > 0.995, f"rollout actions differ from the oracle: {same}"
    np.testing.assert_allclose(g_est[:, :4], est.mean(1)[:, :4], rtol=1e-4, atol=1e-4)
    # replace the device rows by the oracle's so that a (rare) tie-broken action cannot leak into the update check
    tr.t_filled = 0
    tr.load_rows(S, NS, A, R)
    tr.update_round()
    worst = 0.0
    for i in range(5):
        got, want = tr.get_weights(i), agents[i].get_parameters()
        for n in range(len(want)):
            for k in range(6):
                worst = max(worst, float(np.max(np.abs(got[n][k] - want[n][k]) / (1e-5 + 1e-3 * np.abs(want[n][k])))))
    assert worst < 1.0, f"update round differs from the oracle (scaled error {worst})"
    if verbose:
        print(f"smoke ok: rollout action agreement {same:.4f}, update-round scaled error {worst:.3f} (<1), "
              f"{tr.launches} kernel launches on {torch.cuda.get_device_name(0)}")
    return True
