"""The reference-shaped public API (agents/, environments/, training/ + the tensorflow/keras facade in
resilient-consensus-based-marl_b200/) on the GPU, against the golden vectors produced by the reference's own classes
(tests/golden/ref_methods.npz, ref_env.npz) -- the calls below read like the reference's call sites
(training/train_agents.py:105-153, main.py:59-121)."""
import io
import contextlib

import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

from golden_util import load, pretrained          # noqa: E402


def need_gpu():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")


def close(a, b, rtol=1e-4, atol=2e-6):
    np.testing.assert_allclose(np.asarray(a, np.float64), np.asarray(b, np.float64), rtol=rtol, atol=atol)


def build_models(w_agent, NA=5):
    from tensorflow import keras

    def seq(f, n_out, act):
        return keras.Sequential([keras.Input(shape=(NA, f)), keras.layers.Flatten(),
                                 keras.layers.Dense(20, activation=keras.layers.LeakyReLU(alpha=0.1)),
                                 keras.layers.Dense(20, activation=keras.layers.LeakyReLU(alpha=0.1)),
                                 keras.layers.Dense(n_out, activation=act)])
    actor, critic, tr = seq(2, 5, 'softmax'), seq(2, 1, None), seq(3, 1, None)
    actor.set_weights(w_agent[0]); critic.set_weights(w_agent[1]); tr.set_weights(w_agent[2])
    return actor, critic, tr


@pytest.mark.parametrize("H", [0, 1])
def test_rpbcac_agent_methods_match_reference(H):
    need_gpu()
    from agents.resilient_CAC_agents import RPBCAC_agent
    z = load("ref_methods.npz")
    w, _, _ = pretrained()
    s, ns, a, r = z["s"], z["ns"], z["a"], z["r"]
    sa = np.concatenate([s, a], -1)
    actor, critic, tr = build_models(w[0])
    ag = RPBCAC_agent(actor, critic, tr, slow_lr=0.002, fast_lr=0.01, gamma=0.9, H=H)
    cw, closs = ag.critic_update_local(s, ns, r[:, 0])
    tw, tloss = ag.TR_update_local(sa, r[:, 0])
    for k in range(6):
        close(cw[k], z[f"H{H}/critic_local_k{k}"])
        close(tw[k], z[f"H{H}/tr_local_k{k}"])
        assert np.array_equal(critic.get_weights()[k], w[0][1][k])      # own networks untouched (:113,120)
    close(closs, z[f"H{H}/critic_local_loss"], rtol=1e-5)
    close(tloss, z[f"H{H}/tr_local_loss"], rtol=1e-5)
    cmsgs = [cw] + [list(w[j][1]) for j in (1, 2, 4)]                   # device message + plain host lists
    tmsgs = [tw] + [list(w[j][2]) for j in (1, 2, 4)]
    ag.resilient_consensus_critic_hidden(cmsgs)
    ag.resilient_consensus_TR_hidden(tmsgs)
    for k in range(6):
        close(critic.get_weights()[k], z[f"H{H}/critic_after_hidden_k{k}"], rtol=1e-6, atol=1e-6)
        close(tr.get_weights()[k], z[f"H{H}/tr_after_hidden_k{k}"], rtol=1e-6, atol=1e-6)
    cagg = ag.resilient_consensus_critic(s, cmsgs)
    tagg = ag.resilient_consensus_TR(sa, tmsgs)
    close(cagg.numpy(), z[f"H{H}/critic_agg"], rtol=1e-5, atol=3e-6)
    close(tagg.numpy(), z[f"H{H}/tr_agg"], rtol=1e-5, atol=3e-6)
    ag.critic_update_team(s, cagg)
    ag.TR_update_team(sa, tagg)
    for k in range(6):
        close(critic.get_weights()[k], z[f"H{H}/critic_after_team_k{k}"])
        close(tr.get_weights()[k], z[f"H{H}/tr_after_team_k{k}"])
    for step in range(3):                                               # Adam state persists (Appendix A.5)
        al = ag.actor_update(s, ns, sa, a[:, 0])
        close(al, z[f"H{H}/actor_loss_{step}"], rtol=1e-4, atol=1e-6)
        for k in range(6):
            close(actor.get_weights()[k], z[f"H{H}/actor_after_{step}_k{k}"], rtol=2e-4, atol=5e-6)
    close(actor.predict(s[:8]), z[f"H{H}/probs"], rtol=1e-5, atol=1e-6)
    v = critic(s[:1].reshape(1, 5, 2))[0][0].numpy()                    # training/train_agents.py:62
    assert np.ndim(v) == 0 and np.isfinite(v)
    agg = ag._resilient_aggregation(z[f"agg/n4_H{H}_in"])
    close(agg.numpy(), z[f"agg/n4_H{H}_out"], rtol=1e-6, atol=2e-6)
    p = ag.get_parameters()
    assert len(p) == 3 and [x.shape for x in p[1]] == [(10, 20), (20,), (20, 20), (20,), (20, 1), (1,)]


def test_malicious_agent_methods_match_reference():
    need_gpu()
    from agents.adversarial_CAC_agents import Malicious_CAC_agent
    z = load("ref_methods.npz")
    w, _, _ = pretrained()
    s, ns, a, r = z["s"], z["ns"], z["a"], z["r"]
    sa = np.concatenate([s, a], -1)
    actor, critic, tr = build_models(w[4])
    mal = Malicious_CAC_agent(actor, critic, tr, slow_lr=0.002, fast_lr=0.01, gamma=0.9)
    mal.critic_local_weights = w[4][3]                                  # main.py:92
    perms = list(z["mal/perms_96"]) + [z["mal/perm_288"]]
    it = iter(perms)
    mal.perm_source = lambda T: next(it)
    mal.critic_update_local(s, ns, r[:, 4])
    x, xl = mal.TR_update_compromised(sa, -r[:, 0])
    y, yl = mal.critic_update_compromised(s, ns, -r[:, 0])
    big = [np.concatenate([t] * 3, 0) for t in (s, ns, r, a)]
    al = mal.actor_update(big[0], big[1], big[2][:, 4], big[3][:, 4])
    assert next(it, None) is None
    for k in range(6):
        close(mal.critic_local_weights[k], z[f"mal/critic_local_k{k}"], rtol=2e-4, atol=5e-6)
        close(x[k], z[f"mal/tr_k{k}"], rtol=2e-4, atol=5e-6)
        close(y[k], z[f"mal/critic_k{k}"], rtol=2e-4, atol=5e-6)
        close(actor.get_weights()[k], z[f"mal/actor_k{k}"], rtol=2e-4, atol=5e-6)
    close(xl, z["mal/tr_loss"], rtol=1e-4)
    close(yl, z["mal/critic_loss"], rtol=1e-4)
    close(al, z["mal/actor_loss"], rtol=1e-4, atol=1e-6)
    assert len(mal.get_parameters()) == 4


def test_grid_world_api_matches_reference_fixture():
    need_gpu()
    from environments.grid_world import Grid_World
    z = load("ref_env.npz")
    for tag, nrow, na in (("5x5", 5, 5), ("10x10", 10, 16), ("3x3", 3, 3)):
        env = Grid_World(nrow=nrow, ncol=nrow, n_agents=na, desired_state=z[f"{tag}/desired"],
                         initial_state=np.zeros((na, 2), int), randomize_state=True, scaling=True)
        env.state = z[f"{tag}/state_int"][0].copy()
        env._dev_state = None
        for t in range(z[f"{tag}/action"].shape[0]):
            env.step(z[f"{tag}/action"][t])
            st, rw = env.get_data()
            assert np.array_equal(env.state, z[f"{tag}/state_int"][t + 1])
            np.testing.assert_array_equal(st, z[f"{tag}/state_scaled"][t])
            np.testing.assert_allclose(rw, z[f"{tag}/reward_scaled"][t], rtol=1e-7)
    env = Grid_World(nrow=5, ncol=5, n_agents=5, desired_state=z["5x5/desired"], scaling=True, n_envs=7)
    assert env.reset().shape == (7, 5, 2)
    env.step(np.zeros((7, 5)))
    st, rw = env.get_data()
    assert st.shape == (7, 5, 2) and rw.shape == (7, 5)


@pytest.mark.parametrize("n_envs", [1, 16])
def test_train_rpbcac_drop_in_contract(n_envs, tmp_path):
    """What main.py:117-121 does with the return values must keep working."""
    need_gpu()
    from rcmarl import api
    import training.train_agents as training
    w, desired, labels = pretrained()
    cfg = dict(labels=labels, in_nodes=[[0, 1, 2, 3], [1, 2, 3, 4], [2, 3, 4, 0], [3, 4, 0, 1], [4, 0, 1, 2]], weights=w,
               desired=desired, nrow=5, ncol=5, H=1, n_envs=n_envs, gamma=0.9, fast_lr=0.01, slow_lr=0.002, max_ep_len=6,
               n_ep_fixed=8, n_epochs=2, buffer_size=96)
    env, agents, args = api.build_reference_objects(cfg)
    args["n_episodes"] = 20                                             # 2 full blocks + 4 episodes without update
    before = agents[0].critic.get_weights()[0].copy()
    buf = io.StringIO()
    with contextlib.redirect_stdout(buf):
        weights, sim_data = training.train_RPBCAC(env, agents, args)
    lines = [l for l in buf.getvalue().splitlines() if l.startswith("| Episode:")]
    assert len(lines) == 20 and "Est. returns" in lines[0] and "Average actor loss" in lines[0]
    assert list(sim_data.columns) == ["True_team_returns", "True_adv_returns", "Estimated_team_returns"]
    assert len(sim_data) == 20 and np.isfinite(sim_data.to_numpy()).all()
    assert (sim_data["True_team_returns"] <= 0).all()
    sim_data.to_pickle(tmp_path / "sim_data.pkl")                       # main.py:119
    np.save(tmp_path / "pretrained_weights.npy", weights, allow_pickle=True)   # main.py:120
    back = np.load(tmp_path / "pretrained_weights.npy", allow_pickle=True)
    assert back.shape == (5,) and len(back[4]) == 4 and back[0][1][0].shape == (10, 20)
    assert not np.array_equal(agents[0].critic.get_weights()[0], before)       # agents were trained in place
    assert np.array_equal(agents[0].critic.get_weights()[0], back[0][1][0])
    tr = training.train_RPBCAC.last_trainer
    assert tr.t_filled == 96 + 4 * 6 and tr.adam_t[0] == 2 and tr.adam_t[4] == 2


def test_greedy_and_faulty_agent_methods_match_oracle():
    """Per-method API of the two remaining adversaries (agents/adversarial_CAC_agents.py:5-72,184-275) against the
    oracle with injected fit permutations."""
    need_gpu()
    from agents.adversarial_CAC_agents import Greedy_CAC_agent, Faulty_CAC_agent
    from oracle import rpbcac_oracle as O
    z = load("ref_methods.npz")
    w, _, _ = pretrained()
    s, ns, a, r = z["s"], z["ns"], z["a"], z["r"]
    sa = np.concatenate([s, a], -1)
    B = s.shape[0]
    rs = np.random.RandomState(5)
    perms = [rs.permutation(B) for _ in range(21)]
    it = iter(perms)
    g = Greedy_CAC_agent(*build_models(w[3]), slow_lr=0.002, fast_lr=0.01, gamma=0.9)
    g.perm_source = lambda T: next(it)
    x, xl = g.TR_update_local(sa, r[:, 3])
    y, yl = g.critic_update_local(s, ns, r[:, 3])
    al = g.actor_update(s, ns, r[:, 3], a[:, 3])
    og = O.GreedyOracleAgent(w[3][0], w[3][1], w[3][2], 0.002, 0.01, 0.9, dtype=np.float64)
    ox, oxl = og.TR_update_local(sa, r[:, 3], perms[0:10])
    oy, oyl = og.critic_update_local(s, ns, r[:, 3], perms[10:20])
    oal = og.actor_update(s, ns, r[:, 3], a[:, 3], perms[20])
    for k in range(6):
        close(x[k], ox[k], rtol=2e-4, atol=5e-6)
        close(y[k], oy[k], rtol=2e-4, atol=5e-6)
        close(g.actor.get_weights()[k], og.actor[k], rtol=2e-4, atol=5e-6)
    close([xl, yl, al], [oxl, oyl, oal], rtol=1e-4, atol=1e-6)
    f = Faulty_CAC_agent(*build_models(w[2]), slow_lr=0.002, gamma=0.9)
    f.perm_source = lambda T: perms[0]
    before = [m.copy() for m in f.get_critic_weights()]
    fl = f.actor_update(s, ns, r[:, 2], a[:, 2])
    of = O.FaultyOracleAgent(w[2][0], w[2][1], w[2][2], 0.002, 0.9, dtype=np.float64)
    ofl = of.actor_update(s, ns, r[:, 2], a[:, 2], perms[0])
    close(fl, ofl, rtol=1e-4, atol=1e-6)
    for k in range(6):
        close(f.actor.get_weights()[k], of.actor[k], rtol=2e-4, atol=5e-6)
        assert np.array_equal(f.get_critic_weights()[k], before[k]) and np.array_equal(f.get_TR_weights()[k], w[2][2][k])
