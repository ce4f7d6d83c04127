"""Helpers around the public, reference-shaped API (used by bench.py's end-to-end arm and by the tests)."""
import time

import numpy as np
import torch


def build_reference_objects(cfg, n_envs=None):
    """Construct env + agents exactly the way main.py:59-116 does, through the facade and the Agent classes."""
    from tensorflow import keras
    from environments.grid_world import Grid_World
    from agents.resilient_CAC_agents import RPBCAC_agent
    from agents.adversarial_CAC_agents import Faulty_CAC_agent, Greedy_CAC_agent, Malicious_CAC_agent
    labels, w = cfg["labels"], cfg["weights"]
    NA = len(labels)

    def seq(f, n_out, act):
        return keras.Sequential([keras.Input(shape=(NA, f)), keras.layers.Flatten(),
                                 keras.layers.Dense(20, activation=keras.layers.LeakyReLU(alpha=0.1)),
                                 keras.layers.Dense(20, activation=keras.layers.LeakyReLU(alpha=0.1)),
                                 keras.layers.Dense(n_out, activation=act)])
    agents = []
    for i in range(NA):
        actor, critic, tr = seq(2, 5, 'softmax'), seq(2, 1, None), seq(3, 1, None)
        actor.set_weights(w[i][0]); critic.set_weights(w[i][1]); tr.set_weights(w[i][2])
        kw = dict(slow_lr=cfg["slow_lr"], gamma=cfg["gamma"])
        if labels[i] == 'Malicious':
            ag = Malicious_CAC_agent(actor, critic, tr, fast_lr=cfg["fast_lr"], **kw)
            if len(w[i]) > 3:
                ag.critic_local_weights = w[i][3]
        elif labels[i] == 'Faulty':
            ag = Faulty_CAC_agent(actor, critic, tr, **kw)
        elif labels[i] == 'Greedy':
            ag = Greedy_CAC_agent(actor, critic, tr, fast_lr=cfg["fast_lr"], **kw)
        else:
            ag = RPBCAC_agent(actor, critic, tr, fast_lr=cfg["fast_lr"], H=cfg["H"], **kw)
        agents.append(ag)
    env = Grid_World(nrow=cfg["nrow"], ncol=cfg["ncol"], n_agents=NA, desired_state=np.asarray(cfg["desired"]),
                     initial_state=np.zeros((NA, 2), int), randomize_state=True, scaling=True,
                     n_envs=n_envs if n_envs is not None else cfg.get("n_envs", 1))
    args = dict(n_agents=NA, agent_label=list(labels), in_nodes=cfg["in_nodes"], n_actions=5, n_states=2,
                n_episodes=cfg["n_ep_fixed"], max_ep_len=cfg["max_ep_len"], n_ep_fixed=cfg["n_ep_fixed"],
                n_epochs=cfg["n_epochs"], slow_lr=cfg["slow_lr"], fast_lr=cfg["fast_lr"], batch_size=200,
                buffer_size=cfg["buffer_size"], gamma=cfg["gamma"], H=cfg["H"], common_reward=False, random_seed=300)
    return env, agents, args


def export_buffer(tr, keep_rows):
    """Newest `keep_rows` buffer rows as pinned HOST tensors in the reference's exp_buffer layout
    [states (B,NA,2), nstates (B,NA,2), actions (B,NA,1), rewards (B,NA,1)] (train_agents.py:36-40)."""
    B = tr.t_filled * tr.N
    lo = max(0, B - keep_rows)
    sa = tr.sa[lo:B].view(-1, tr.NA, 3)
    host = [sa[:, :, :2].contiguous().cpu().pin_memory(), tr.ns[lo:B].view(-1, tr.NA, 2).cpu().pin_memory(),
            sa[:, :, 2:].contiguous().cpu().pin_memory(), tr.r[lo:B].view(-1, tr.NA, 1).cpu().pin_memory()]
    return host


def timed_train(cfg, host_buffer, n_blocks, rank=0, world=1):
    """One call of the public training function on host inputs, wall-clock timed between device synchronisations."""
    import training.train_agents as training
    env, agents, args = build_reference_objects(cfg)
    args["n_episodes"] = cfg["n_ep_fixed"] * n_blocks
    torch.cuda.synchronize()
    if world > 1:
        import torch.distributed as dist
        dist.barrier()
    t0 = time.perf_counter()
    weights, sim = training.train_RPBCAC(env, agents, args, exp_buffer=host_buffer, verbose=False)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    tr = training.train_RPBCAC.last_trainer
    if world > 1:
        import torch.distributed as dist
        t = torch.tensor([dt], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dt = float(t.item())
    return dict(seconds=dt, h2d_bytes=int(tr.h2d_bytes), d2h_bytes=int(tr.d2h_bytes), weights=weights, sim=sim)
