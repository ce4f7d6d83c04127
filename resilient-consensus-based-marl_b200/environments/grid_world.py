"""Grid_World with the reference's interface (environments/grid_world.py:5-75), batched over n_envs independent
environments whose state lives in device memory; transitions run in the env_step kernel of librcmarl.so.

n_envs == 1 (the default, as constructed by main.py:109-116) returns exactly the reference's shapes.  The reset
draws come from NumPy's global RNG in the reference's call (`randint([0,0],[nrow,ncol],size=(n_agents,2))`), so a
seeded single-environment run visits the same initial states as the reference.

Reproduced quirk: the collision test of the reference (:56) includes the agent itself, so the
"moves to a new cell" branch (:59-60) is dead: reward = 0 if the agent is on its goal and stays, else
-(pre-move L1 distance to the goal) - 1; both coordinates are clipped with nrow (:55)."""
import os

import numpy as np
import gym
from gym import spaces  # noqa: F401


class Grid_World(gym.Env):
    metadata = {'render.modes': ['console']}

    def __init__(self, nrow=5, ncol=5, n_agents=1, desired_state=None, initial_state=None, randomize_state=True,
                 scaling=False, n_envs=None):
        self.nrow = nrow
        self.ncol = ncol
        self.n_agents = n_agents
        self.initial_state = initial_state
        self.desired_state = desired_state
        self.randomize_state = randomize_state
        self.n_states = 2
        self.n_envs = int(n_envs if n_envs is not None else os.environ.get("RCMARL_N_ENVS", 1))
        self.actions_dict = {0: np.array([0, 0]), 1: np.array([-1, 0]), 2: np.array([1, 0]), 3: np.array([0, -1]),
                             4: np.array([0, 1])}
        self._dev_state = None
        self._dev_desired = None
        self.reset()
        if scaling:
            x, y = np.arange(nrow), np.arange(ncol)
            self.mean_state = np.array([np.mean(x), np.mean(y)])
            self.std_state = np.array([np.std(x), np.std(y)])
        else:
            self.mean_state, self.std_state = 0, 1

    def _squeeze(self, a):
        return a[0] if self.n_envs == 1 else a

    def reset(self):
        '''Resets the environment(s)'''
        shape = (self.n_agents, self.n_states) if self.n_envs == 1 else (self.n_envs, self.n_agents, self.n_states)
        if self.randomize_state:
            self.state = np.random.randint([0, 0], [self.nrow, self.ncol], size=shape)
        else:
            self.state = np.broadcast_to(np.array(self.initial_state), shape).copy()
        self.reward = np.zeros(shape[:-1])
        self._dev_state = None
        return self.state

    def step(self, action):
        '''Transition + rewards for every environment in one kernel launch.'''
        import torch
        from rcmarl import ops
        dev = torch.device("cuda", torch.cuda.current_device())
        if self._dev_state is None:
            st = np.asarray(self.state, np.int32).reshape(self.n_envs, self.n_agents, 2)
            self._dev_state = torch.as_tensor(st).to(dev).contiguous()
        if self._dev_desired is None:
            self._dev_desired = torch.as_tensor(np.asarray(self.desired_state, np.int32).reshape(self.n_agents, 2)).to(dev)
        act = ops.dev_f32(np.asarray(action, np.float32).reshape(self.n_envs, self.n_agents))
        rew = ops.env_step(self._dev_state, act, self._dev_desired, self.nrow)
        self.state = self._squeeze(self._dev_state.cpu().numpy().astype(np.int64))
        # the kernel returns reward/5 rounded to float32 (what the training tensors hold, train_agents.py:91)
        self._reward_scaled = self._squeeze(rew.cpu().numpy().astype(np.float64))
        self.reward = self._reward_scaled * 5

    def get_data(self):
        '''Returns scaled state and scaled reward (grid_world.py:66-72)'''
        state_scaled = (self.state - self.mean_state) / self.std_state
        reward_scaled = getattr(self, "_reward_scaled", None)
        if reward_scaled is None or self._dev_state is None:
            reward_scaled = self.reward / 5
        return state_scaled, reward_scaled

    def close(self):
        pass
