"""The NumPy restatement (oracle/rpbcac_oracle.py) against the reference sources
executed verbatim on oracle/tf_facade (fixtures: tests/golden/ref_*.npz, made by
oracle/make_golden.py).  fp32 tolerances: forward 1e-5/1e-6, weights after a fit
or projection step 1e-4/1e-6."""
import numpy as np
import pytest

from golden_util import load, agent_weights, pretrained
from oracle import rpbcac_oracle as O


def close(a, b, rtol=1e-4, atol=2e-6):
    np.testing.assert_allclose(np.asarray(a, np.float64), np.asarray(b, np.float64), rtol=rtol, atol=atol)


def test_env_matches_reference():
    z = load("ref_env.npz")
    for tag, nrow, na in (("5x5", 5, 5), ("10x10", 10, 16), ("3x3", 3, 3)):
        env = O.GridWorldOracle(nrow, nrow, na, z[f"{tag}/desired"], n_envs=1)
        S = z[f"{tag}/state_int"]
        env.set_state(S[0])
        for t in range(z[f"{tag}/action"].shape[0]):
            env.step(z[f"{tag}/action"][t])
            st, rw = env.get_data()
            assert np.array_equal(env.state[0], S[t + 1])
            np.testing.assert_array_equal(st[0], z[f"{tag}/state_scaled"][t])
            np.testing.assert_array_equal(rw[0], z[f"{tag}/reward_scaled"][t])


def test_aggregation_matches_reference():
    z = load("ref_methods.npz")
    for n, H in ((4, 0), (4, 1), (6, 2), (9, 4), (3, 1)):
        got = O.resilient_aggregation(z[f"agg/n{n}_H{H}_in"], H)
        close(got, z[f"agg/n{n}_H{H}_out"], rtol=1e-6, atol=1e-6)


@pytest.mark.parametrize("H", [0, 1])
def test_cooperative_methods_match_reference(H):
    z = load("ref_methods.npz")
    w, _, _ = pretrained()
    s, ns, a, r = z["s"], z["ns"], z["a"], z["r"]
    sa = np.concatenate([s, a], -1)
    ag = O.RPBCACOracleAgent(w[0][0], w[0][1], w[0][2], 0.002, 0.01, gamma=0.9, H=H)
    assert bool(z[f"H{H}/critic_after_local_same"])
    cw, closs = ag.critic_update_local(s, ns, r[:, 0])
    tw, tloss = ag.TR_update_local(sa, r[:, 0])
    for k in range(6):
        close(cw[k], z[f"H{H}/critic_local_k{k}"])
        close(tw[k], z[f"H{H}/tr_local_k{k}"])
        assert np.array_equal(ag.critic[k], w[0][1][k])          # internal weights restored
    close(closs, z[f"H{H}/critic_local_loss"], rtol=1e-5)
    close(tloss, z[f"H{H}/tr_local_loss"], rtol=1e-5)
    cm = [cw] + [w[j][1] for j in (1, 2, 4)]
    tm = [tw] + [w[j][2] for j in (1, 2, 4)]
    ag.resilient_consensus_critic_hidden(cm)
    ag.resilient_consensus_TR_hidden(tm)
    for k in range(6):
        close(ag.critic[k], z[f"H{H}/critic_after_hidden_k{k}"], rtol=1e-6, atol=1e-7)
        close(ag.TR[k], z[f"H{H}/tr_after_hidden_k{k}"], rtol=1e-6, atol=1e-7)
    cagg = ag.resilient_consensus_critic(s, cm)
    tagg = ag.resilient_consensus_TR(sa, tm)
    close(cagg, z[f"H{H}/critic_agg"], rtol=1e-5, atol=2e-6)
    close(tagg, z[f"H{H}/tr_agg"], rtol=1e-5, atol=2e-6)
    ag.critic_update_team(s, cagg)
    ag.TR_update_team(sa, tagg)
    for k in range(6):
        close(ag.critic[k], z[f"H{H}/critic_after_team_k{k}"])
        close(ag.TR[k], z[f"H{H}/tr_after_team_k{k}"])
    for step in range(3):
        al = ag.actor_update(s, ns, sa, a[:, 0])
        close(al, z[f"H{H}/actor_loss_{step}"], rtol=1e-4, atol=1e-6)
        for k in range(6):
            close(ag.actor[k], z[f"H{H}/actor_after_{step}_k{k}"], rtol=2e-4, atol=5e-6)
    close(ag.action_probs(s[:8]), z[f"H{H}/probs"], rtol=1e-5, atol=1e-6)


def test_malicious_methods_match_reference():
    z = load("ref_methods.npz")
    w, _, _ = pretrained()
    s, ns, a, r = z["s"], z["ns"], z["a"], z["r"]
    sa = np.concatenate([s, a], -1)
    perms = z["mal/perms_96"]
    mal = O.MaliciousOracleAgent(w[4][0], w[4][1], w[4][2], 0.002, 0.01, gamma=0.9, critic_local_w=w[4][3])
    mal.critic_update_local(s, ns, r[:, 4], perms[0:10])
    x, xl = mal.TR_update_compromised(sa, -r[:, 0], perms[10:20])
    y, yl = mal.critic_update_compromised(s, ns, -r[:, 0], perms[20:30])
    big = [np.concatenate([t] * 3, 0) for t in (s, ns, r, a)]
    al = mal.actor_update(big[0], big[1], big[2][:, 4], big[3][:, 4], z["mal/perm_288"])
    for k in range(6):
        close(mal.critic_local_weights[k], z[f"mal/critic_local_k{k}"], rtol=2e-4, atol=5e-6)
        close(x[k], z[f"mal/tr_k{k}"], rtol=2e-4, atol=5e-6)
        close(y[k], z[f"mal/critic_k{k}"], rtol=2e-4, atol=5e-6)
        close(mal.actor[k], z[f"mal/actor_k{k}"], rtol=2e-4, atol=5e-6)
    close(xl, z["mal/tr_loss"], rtol=1e-4)
    close(yl, z["mal/critic_loss"], rtol=1e-4)
    close(al, z["mal/actor_loss"], rtol=1e-4, atol=1e-6)


def test_update_rounds_match_reference_train_run():
    """Two update rounds of reference train_RPBCAC (verbatim) replayed through
    oracle.update_round from the recorded buffer + permutations."""
    z = load("ref_train_run.npz")
    w, desired, labels = pretrained()
    assert np.array_equal(desired, z["desired"])
    in_nodes = [[0, 1, 2, 3], [1, 2, 3, 4], [2, 3, 4, 0], [3, 4, 0, 1], [4, 0, 1, 2]]
    agents = []
    for i in range(5):
        if labels[i] == "Malicious":
            agents.append(O.MaliciousOracleAgent(w[i][0], w[i][1], w[i][2], 0.002, 0.01, 0.9, critic_local_w=w[i][3]))
        else:
            agents.append(O.RPBCACOracleAgent(w[i][0], w[i][1], w[i][2], 0.002, 0.01, 0.9, H=1))
    perms = [z[f"perm{j}"] for j in range(int(z["n_perms"]))]
    it = iter(perms)

    def perm_source(T):
        p = next(it)
        assert len(p) == T
        return p
    rows_per_round = 250
    for rnd in (1, 2):
        B = rows_per_round * rnd
        O.update_round(agents, labels, in_nodes, z["s"][:B], z["ns"][:B], z["a"][:B], z["r"][:B],
                       n_envs=1, n_epochs=2, n_actor_steps=rows_per_round, common_reward=False,
                       perm_source=perm_source)
    assert next(it, None) is None                                # every injected permutation consumed
    final = agent_weights(z, "final")
    for i in range(5):
        got = agents[i].get_parameters()
        for n in range(len(final[i])):
            for k in range(6):
                close(got[n][k], final[i][n][k], rtol=5e-4, atol=2e-5)


def test_greedy_methods_match_reference():
    """Greedy_CAC_agent (agents/adversarial_CAC_agents.py:184-275) executed verbatim on the facade vs the oracle class."""
    z = load("ref_adversaries.npz")
    w, _, _ = pretrained()
    s, ns, a, r = z["s"], z["ns"], z["a"], z["r"]
    sa = np.concatenate([s, a], -1)
    perms = z["greedy/perms_96"]
    gr = O.GreedyOracleAgent(w[3][0], w[3][1], w[3][2], 0.002, 0.01, gamma=0.9)
    x, xl = gr.TR_update_local(sa, r[:, 3], perms[0:10])
    y, yl = gr.critic_update_local(s, ns, r[:, 3], perms[10:20])
    big = [np.concatenate([t] * 3, 0) for t in (s, ns, r, a)]
    al = gr.actor_update(big[0], big[1], big[2][:, 3], big[3][:, 3], z["greedy/perm_288"])
    for k in range(6):
        close(x[k], z[f"greedy/tr_k{k}"], rtol=2e-4, atol=5e-6)
        close(y[k], z[f"greedy/critic_k{k}"], rtol=2e-4, atol=5e-6)
        close(gr.actor[k], z[f"greedy/actor_k{k}"], rtol=2e-4, atol=5e-6)
    close(xl, z["greedy/tr_loss"], rtol=1e-4)
    close(yl, z["greedy/critic_loss"], rtol=1e-4)
    close(al, z["greedy/actor_loss"], rtol=1e-4, atol=1e-6)
    assert len(gr.get_parameters()) == int(z["greedy/n_param_lists"]) == 3


def test_faulty_methods_match_reference():
    """Faulty_CAC_agent (agents/adversarial_CAC_agents.py:5-72): only the actor learns, critic / TR are transmitted as is."""
    z = load("ref_adversaries.npz")
    w, _, _ = pretrained()
    s, ns, a, r = z["s"], z["ns"], z["a"], z["r"]
    assert bool(z["faulty/critic_unchanged"]) and bool(z["faulty/tr_unchanged"])
    fa = O.FaultyOracleAgent(w[4][0], w[4][1], w[4][2], 0.002, gamma=0.9)
    big = [np.concatenate([t] * 3, 0) for t in (s, ns, r, a)]
    al = fa.actor_update(big[0], big[1], big[2][:, 4], big[3][:, 4], z["faulty/perm_288"])
    close(al, z["faulty/actor_loss"], rtol=1e-4, atol=1e-6)
    for k in range(6):
        close(fa.actor[k], z[f"faulty/actor_k{k}"], rtol=2e-4, atol=5e-6)
        assert np.array_equal(fa.get_critic_weights()[k], z[f"faulty/critic_k{k}"])
        assert np.array_equal(fa.get_TR_weights()[k], z[f"faulty/tr_k{k}"])


def test_update_rounds_with_greedy_and_faulty_match_reference_train_run():
    """Reference train_RPBCAC verbatim with 3 cooperative + 1 greedy + 1 faulty agent and common_reward=True
    (train_agents.py:106; the *_global scenarios), replayed through oracle.update_round."""
    z = load("ref_adversaries.npz")
    w, desired, _ = pretrained()
    labels = [str(x) for x in z["run/labels"]]
    assert np.array_equal(desired, z["run/desired"])
    in_nodes = [[0, 1, 2, 3], [1, 2, 3, 4], [2, 3, 4, 0], [3, 4, 0, 1], [4, 0, 1, 2]]
    agents = []
    for i, l in enumerate(labels):
        if l == "Greedy":
            agents.append(O.GreedyOracleAgent(w[i][0], w[i][1], w[i][2], 0.002, 0.01, 0.9))
        elif l == "Faulty":
            agents.append(O.FaultyOracleAgent(w[i][0], w[i][1], w[i][2], 0.002, 0.9))
        else:
            agents.append(O.RPBCACOracleAgent(w[i][0], w[i][1], w[i][2], 0.002, 0.01, 0.9, H=1))
    perms = [z[f"run/perm{j}"] for j in range(int(z["run/n_perms"]))]
    it = iter(perms)

    def perm_source(T):
        p = next(it)
        assert len(p) == T
        return p
    for rnd in (1, 2):
        B = 250 * rnd
        O.update_round(agents, labels, in_nodes, z["run/s"][:B], z["run/ns"][:B], z["run/a"][:B], z["run/r"][:B],
                       n_envs=1, n_epochs=2, n_actor_steps=250, common_reward=True, perm_source=perm_source)
    assert next(it, None) is None
    final = agent_weights(z, "run/final")
    for i in range(5):
        got = agents[i].get_parameters()
        assert len(got) == len(final[i]) == 3
        for n in range(3):
            for k in range(6):
                close(got[n][k], final[i][n][k], rtol=5e-4, atol=2e-5)
