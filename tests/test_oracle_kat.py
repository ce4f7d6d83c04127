"""Known-answer test: the oracle's critic forward + env reset + state scaling +
NumPy RNG draw order reproduce the `Est. returns` logged by the reference's real
TensorFlow runs (fixtures from oracle/make_golden.py)."""
import numpy as np
import pytest

from golden_util import load, agent_weights
from oracle import rpbcac_oracle as O

TAGS = ["malicious_H1_s100", "coop_H0_s200", "greedy_H1_s300", "faulty_global_H1_s100"]


@pytest.mark.parametrize("tag", TAGS)
def test_est_returns_match_tf_log(tag):
    z = load("kat_est_returns.npz")
    w = agent_weights(z, tag)
    labels = [str(x) for x in z[f"{tag}/labels"]]
    seed = int(z[f"{tag}/seed"])
    expect = z[f"{tag}/est_returns"]                        # (50, n_coop)
    coop = [i for i, l in enumerate(labels) if l == "Cooperative"]
    np.random.seed(seed)                                     # main.py:46
    s_desired = np.random.randint(0, 5, size=(5, 2))         # main.py:48
    _ = np.random.randint(0, 5, size=(5, 2))                 # main.py:49
    assert np.array_equal(s_desired, z[f"{tag}/desired"])
    env = O.GridWorldOracle(5, 5, 5, s_desired, n_envs=1)
    env.reset_np_global()                                    # Grid_World.__init__ -> reset (grid_world.py:28)
    p = np.full(5, 0.2)
    worst = 0.0
    for t in range(expect.shape[0]):
        env.reset_np_global()                                # train_agents.py:55
        state, _ = env.get_data()
        x = O.flatten_rows(state, np.float32)
        got = [float(O.mlp_forward(O.cast_weights(w[i][1], np.float32), x)[0, 0]) for i in coop]
        worst = max(worst, np.abs(np.array(got) - expect[t]).max())
        for _j in range(20):                                 # consume the episode's draws (a3)
            for _node in range(5):
                np.random.choice(5)
                np.random.choice(5, p=p)
                np.random.choice([0, 1], p=[0.9, 0.1])
    assert worst < 5e-6, worst


def test_choice_equivalences():
    """np.random.choice(n, p) == searchsorted(cumsum(p)/sum, U, 'right') and
    choice([a,b],[1-mu,mu]) == a if U < 1-mu else b  ('RNG')."""
    rs = np.random.RandomState(0)
    for seed in range(200):
        p = rs.dirichlet(np.ones(5)).astype(np.float32)
        p = p / p.sum()
        np.random.seed(seed)
        c1 = np.random.choice(5, p=p)
        c2 = np.random.choice([7, 9], p=[0.9, 0.1])
        np.random.seed(seed)
        u1 = np.random.random_sample()
        u2 = np.random.random_sample()
        u = np.array([0.0, u1, u2], np.float32)
        got = O.sample_action_from_uniforms(p, u, mu=0.1)
        cdf = np.cumsum(p.astype(np.float64))
        cdf /= cdf[-1]
        assert c1 == np.searchsorted(cdf, u1, side="right")
        assert (c2 == 7) == (u2 < 0.9)
        if abs(cdf - u1).min() > 1e-6 and abs(u2 - 0.9) > 1e-6:   # away from fp32 rounding ties
            assert got == (c1 if c2 == 7 else 0)
