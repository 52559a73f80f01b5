"""A deterministic stand-in for the reference's EnvironmentHelper (src/environments/helper.py:12-67,
running_gym_sequential_vectorized.py:19-100): numpy dynamics on the CPU, float64 observations and rewards like gym."""
from types import SimpleNamespace

import numpy as np
import torch


class FakeHelper:

    def __init__(self, run, n_envs, obs_dim, act_dim, seed=0):
        self.run = run
        self.n, self.d, self.a = n_envs, obs_dim, act_dim
        rng = np.random.default_rng(seed)
        self.A = rng.standard_normal((obs_dim, obs_dim)) * (0.5 / np.sqrt(obs_dim))
        self.B = rng.standard_normal((obs_dim, act_dim)) * 0.3
        self.x0 = rng.standard_normal((n_envs, obs_dim))
        self.timestep = SimpleNamespace(observation=None, reward=None, terminated=None, truncated=None)
        self.memory = []
        self.t = 0

    def reset(self, release_memory=True):
        self.memory = []

    def reset_environment(self, test_phase=False):
        self.x = self.x0.copy()
        self.t = 0
        self.timestep.terminated = np.zeros(self.n, dtype=np.bool_)
        self.timestep.truncated = np.zeros(self.n, dtype=np.bool_)

    def get_state(self, test_phase=False):
        return torch.tensor(self.x).to(torch.float32)[:, None, :]  # [N, window=1, obs]

    def step(self, action):
        a = action.detach().cpu().double().numpy() if isinstance(action, torch.Tensor) else np.asarray(action, dtype=np.float64)
        self.x = np.tanh(self.x @ self.A.T + a @ self.B.T)
        self.t += 1
        self.timestep.reward = self.x.sum(axis=1) * 0.1 - (a ** 2).sum(axis=1) * 0.01
        self.timestep.terminated = ((np.arange(self.n) + self.t) % 11 == 0)
        self.timestep.truncated = np.zeros(self.n, dtype=np.bool_)
        self.x = np.where(self.timestep.terminated[:, None], self.x0, self.x)  # terminated envs restart
