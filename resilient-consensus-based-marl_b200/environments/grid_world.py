"""Grid_World with the reference's interface (environments/grid_world.py:5-75 of
mfigura/Resilient-consensus-based-MARL), batched over `n_envs` independent environments whose integer positions live
in device memory; transitions and rewards are computed by the `env_step` kernel of librcmarl.so.

`n_envs == 1` (the default, as constructed by main.py:109-116) returns exactly the reference's shapes.  Reset draws
come from NumPy's global RNG with the reference's call, so a seeded single-environment run visits the same initial
states as the reference.

Reproduced quirk: the reference's collision test (:56) includes the agent itself, so its "moved to a
free cell" branch (:59-60) is dead: reward = 0 if the agent sits on its goal and stays, else -(pre-move L1 distance
to the goal) - 1; both coordinates are clipped with `nrow` (:55); get_data scales the reward by 1/5 (:71).
"""
import os

import numpy as np
import gym

# action id -> (d_row, d_col): stay, up, down, left, right (grid_world.py:27)
MOVES = ((0, 0), (-1, 0), (1, 0), (0, -1), (0, 1))


class Grid_World(gym.Env):
    metadata = {'render.modes': ['console']}

    def __init__(self, nrow=5, ncol=5, n_agents=1, desired_state=None, initial_state=None, randomize_state=True,
                 scaling=False, n_envs=None):
        if n_envs is None:
            n_envs = int(os.environ.get("RCMARL_N_ENVS", 1))
        self.__dict__.update(nrow=nrow, ncol=ncol, n_agents=n_agents, n_states=2, n_envs=int(n_envs),
                             desired_state=desired_state, initial_state=initial_state,
                             randomize_state=randomize_state)
        self.actions_dict = {a: np.array(m) for a, m in enumerate(MOVES)}
        axes = (np.arange(nrow), np.arange(ncol))
        # state scaling statistics (grid_world.py:30-35); identity when scaling is off
        self.mean_state = np.array([ax.mean() for ax in axes]) if scaling else 0
        self.std_state = np.array([ax.std() for ax in axes]) if scaling else 1
        self._dev_state = self._dev_desired = self._reward_scaled = None
        self.reset()

    # ------------------------------------------------------------------ helpers
    @property
    def _shape(self):
        head = () if self.n_envs == 1 else (self.n_envs,)
        return head + (self.n_agents, self.n_states)

    def _per_env(self, a):
        return a[0] if self.n_envs == 1 else a

    def _device_state(self):
        import torch
        dev = torch.device("cuda", torch.cuda.current_device())
        if self._dev_state is None:
            host = np.asarray(self.state, np.int32).reshape(self.n_envs, self.n_agents, 2)
            self._dev_state = torch.as_tensor(host).to(dev).contiguous()
        if self._dev_desired is None:
            goal = np.asarray(self.desired_state, np.int32).reshape(self.n_agents, 2)
            self._dev_desired = torch.as_tensor(goal).to(dev)
        return self._dev_state, self._dev_desired

    # ------------------------------------------------------------------ reference API
    def reset(self):
        """New initial positions for every environment (grid_world.py:37-45)."""
        if self.randomize_state:
            self.state = np.random.randint([0, 0], [self.nrow, self.ncol], size=self._shape)
        else:
            self.state = np.array(np.broadcast_to(np.asarray(self.initial_state), self._shape))
        self.reward = np.zeros(self._shape[:-1])
        self._dev_state = self._reward_scaled = None
        return self.state

    def step(self, action):
        """One transition + rewards of all environments in a single kernel launch (grid_world.py:47-64)."""
        from rcmarl import ops
        state, goal = self._device_state()
        act = ops.dev_f32(np.asarray(action, np.float32).reshape(self.n_envs, self.n_agents))
        scaled = ops.env_step(state, act, goal, self.nrow)          # reward / 5 in float32, as the training tensors hold it
        self.state = self._per_env(state.cpu().numpy().astype(np.int64))
        self._reward_scaled = self._per_env(scaled.cpu().numpy().astype(np.float64))
        self.reward = 5 * self._reward_scaled

    def get_data(self):
        """(scaled state, scaled reward) as float64 arrays (grid_world.py:66-72)."""
        scaled_reward = self.reward / 5 if self._reward_scaled is None else self._reward_scaled
        return (self.state - self.mean_state) / self.std_state, scaled_reward

    def close(self):
        pass
