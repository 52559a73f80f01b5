"""CPU restatement of the reference's Soft Actor-Critic update step — TEST INFRASTRUCTURE ONLY.

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s CPU legs may import this module; the product path
(`mujoco_reinforcement_learning_b200/`) never does.

Follows, line by line:
  * `SoftActorCritic.train`   src/entities/algorithms/soft_actor_critic.py:33-118
  * `soft_update` / `hard_update`   soft_actor_critic.py:12-19
  * `QNetwork` (twin MLPs over cat[state, action])   src/models/linear/q_network.py:7-38
  * `Agent.act` with `Normal.rsample()`   src/entities/agents/agent.py:26-42
  * `SoftActorCriticAgent.initialize_networks`   src/entities/agents/soft_actor_critic_agent.py:12-35
    (with the MLP `linear.Actor` / `linear.QNetwork`; as committed the agent binds the Transformer variants)
The two sources of randomness — `torch.randperm` (:39) and the standard-normal draws inside `rsample()` (:52, :78) —
are inputs, so that a CUDA implementation can be fed the same numbers.
Pinned by `tests/golden/ref_sac_*.npz`, written by running the reference's own `train` (oracle/make_golden.py).
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Dict, Optional, Sequence, Tuple

import torch
from torch import nn
from torch.nn.functional import mse_loss

from oracle.ppo_oracle import MLPBlock, OracleActor


class OracleQNetwork(nn.Module):
    """src/models/linear/q_network.py:7-38 — two independent MLPs on cat[state.flatten(1), action]."""

    def __init__(self, in_dim: int, hidden: Sequence[int], activation: str):
        super().__init__()
        self.first_network = MLPBlock(in_dim, hidden, 1, activation, final_tanh=False)
        self.second_network = MLPBlock(in_dim, hidden, 1, activation, final_tanh=False)

    def forward(self, state, action):
        x = torch.cat([state.reshape(len(state), -1), action], 1)
        return self.first_network(x), self.second_network(x)


@dataclass
class SacConfig:
    state_dim: int            # flattened state width (input_shape * window_length; q_network.py:19 assumes window 2)
    act_dim: int
    hidden: Sequence[int]
    activation: str = "tanh"
    output_max_value: float = 1.0
    learning_rate: float = 1e-4
    batch_size: int = 64
    gamma: float = 0.99
    alpha: float = 0.05
    tau: float = 0.005
    target_update_interval: int = 1
    max_grad_norm: float = 1.0  # Run.instance().ppo_config.max_grad_norm (:67, :84)


class OracleSacAgent:
    """soft_actor_critic_agent.py:12-35 (construction order = order of random draws)."""

    def __init__(self, cfg: SacConfig):
        self.cfg = cfg
        self.networks = nn.ModuleDict()
        self.networks["actor"] = OracleActor(cfg.state_dim, cfg.act_dim, cfg.hidden, cfg.activation, cfg.output_max_value)
        self.networks["online_critic"] = OracleQNetwork(cfg.state_dim + cfg.act_dim, cfg.hidden, cfg.activation)
        self.networks["target_critic"] = OracleQNetwork(cfg.state_dim + cfg.act_dim, cfg.hidden, cfg.activation)
        self.optimizers = {
            "actor": torch.optim.Adam(self.networks["actor"].parameters(), lr=cfg.learning_rate, foreach=False),
            "online_critic": torch.optim.Adam(self.networks["online_critic"].parameters(), lr=cfg.learning_rate, foreach=False),
        }
        hard_update(self.networks["target_critic"], self.networks["online_critic"])  # SoftActorCritic.__init__, :30


def soft_update(target: nn.Module, source: nn.Module, tau: float):
    for tp, p in zip(target.parameters(), source.parameters()):  # :12-14
        tp.data.copy_(tp.data * (1.0 - tau) + p.data * tau)


def hard_update(target: nn.Module, source: nn.Module):
    for tp, p in zip(target.parameters(), source.parameters()):  # :17-19
        tp.data.copy_(p.data)


def _log_prob_sum(mean, std, value):
    """`distributions.log_prob(x).sum(dim=1)[:, None]` (:22-23) with torch's Normal.log_prob formula."""
    var = std ** 2
    return (-((value - mean) ** 2) / (2 * var) - std.log() - math.log(math.sqrt(2 * math.pi))).sum(dim=1)[:, None]


def sac_train_step(agent: OracleSacAgent, memory: Dict[str, torch.Tensor], idx: torch.Tensor, eps_next: torch.Tensor,
                   eps_pi: torch.Tensor, update_count: int) -> Tuple[float, float, float, float, float]:
    """One call of `SoftActorCritic.train` (:33-118), automatic entropy tuning off.

    memory: flattened leaves `current_state [M, ...]`, `next_state [M, ...]`, `reward [M,1]`, `action [M,A]`,
    `is_alive [M,1]` (bool).  idx: the permutation of :39.  eps_next / eps_pi [batch, A]: the standard-normal draws of
    the two `rsample()` calls (:52 and :78).  Returns (qf1_loss, qf2_loss, policy_loss, mean min_qf_pi, alpha_loss=0).
    """
    cfg = agent.cfg
    B = cfg.batch_size
    shuffled = {k: v.clone()[idx] for k, v in memory.items()}                       # :40
    shuffled["reward"] = shuffled["reward"] - shuffled["reward"].mean()             # :41
    shuffled["reward"] = shuffled["reward"] / shuffled["reward"].std()              # :42 (unbiased std over everything)
    batch = {k: v[0:B] for k, v in shuffled.items()}                                # :44 (batches_per_timestep = 1)
    s, s2, r, a, mask = batch["current_state"], batch["next_state"], batch["reward"], batch["action"], batch["is_alive"]
    actor, online, target = agent.networks["actor"], agent.networks["online_critic"], agent.networks["target_critic"]
    with torch.no_grad():                                                           # :50-59
        mean2, std2 = actor(s2)
        a2 = mean2 + eps_next * std2
        logp2 = _log_prob_sum(mean2, std2, a2)
        q1t, q2t = target(s2, a2)
        min_q = torch.min(q1t, q2t) - cfg.alpha * logp2
        y = (r + mask * cfg.gamma * min_q).to(torch.float32)
    q1, q2 = online(s, a)                                                           # :60-62
    qf1_loss, qf2_loss = mse_loss(q1, y), mse_loss(q2, y)                           # :63-68
    qf_loss = qf1_loss + qf2_loss
    agent.optimizers["online_critic"].zero_grad()
    qf_loss.backward()
    torch.nn.utils.clip_grad_norm_(online.parameters(), cfg.max_grad_norm)          # :73-74 (before the step: effective)
    agent.optimizers["online_critic"].step()
    mean, std = actor(s)                                                            # :77 rsample
    a_pi = mean + eps_pi * std
    q1p, q2p = online(s, a_pi)                                                      # :79-80 (updated critic)
    logp = _log_prob_sum(mean, std, a_pi)
    min_q_pi = torch.min(q1p, q2p)
    policy_loss = (cfg.alpha * logp - min_q_pi).mean()                              # :84-86
    agent.optimizers["actor"].zero_grad()
    policy_loss.backward()
    torch.nn.utils.clip_grad_norm_(actor.parameters(), cfg.max_grad_norm)           # :90-91
    agent.optimizers["actor"].step()
    if update_count % cfg.target_update_interval == 0:                              # :109-111
        soft_update(target, online, cfg.tau)
    return (qf1_loss.item(), qf2_loss.item(), policy_loss.item(), min_q_pi.mean().item(), 0.0)


def synthetic_replay(n_envs: int, steps: int, state_shape: Sequence[int], act_dim: int, seed: int) -> Dict[str, torch.Tensor]:
    """Seeded replay memory `[N, T, ...]` with the leaves of soft_actor_critic.py:151-168."""
    g = torch.Generator().manual_seed(seed)
    s = torch.randn(n_envs, steps, *state_shape, generator=g)
    return {"current_state": s, "next_state": s + 0.1 * torch.randn(n_envs, steps, *state_shape, generator=g),
            "action": torch.randn(n_envs, steps, act_dim, generator=g).clamp_(-3, 3),
            "reward": torch.randn(n_envs, steps, 1, generator=g) * 2.0 + 0.5,
            "is_alive": torch.rand(n_envs, steps, 1, generator=g) > 0.05}
