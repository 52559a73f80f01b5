"""Torch-CPU restatement of the reference PPO hot path — TEST INFRASTRUCTURE ONLY.

The reference computes everything with stock torch CPU operators, so the faithful CPU
restatement is written with the same operators (`nn.Linear`, `Normal`, `huber_loss`,
`optim.Adam`).  Every function cites the reference lines it follows (paths relative to
`/root/reference/`).  `oracle/naive.py` is the independent float64 check of this file,
and `tests/golden/ref_*.npz` (written by `oracle/make_golden.py` from the reference's
own code) pins it.

Parity status: see `oracle/__init__.py` — everything pinned by reference outputs
except the torchrl GAE recurrence, which is **parity unpinned** (third-party
`torchrl==0.6.0`, absent from the tree; restated from its published algorithm).
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence, Tuple

import torch
from torch import nn
from torch.nn.functional import huber_loss


# ----------------------------------------------------------------------------------
# GAE — torchrl==0.6.0 objectives/value/functional.py::generalized_advantage_estimate
# (third-party, pinned in requirements.txt:7; call site src/entities/algorithms/ppo.py:76-80)
# ----------------------------------------------------------------------------------
def generalized_advantage_estimate(gamma: float,
                                   lmbda: float,
                                   state_value: torch.Tensor,
                                   next_state_value: torch.Tensor,
                                   reward: torch.Tensor,
                                   done: torch.Tensor,
                                   terminated: Optional[torch.Tensor] = None,
                                   time_dim: int = -2) -> Tuple[torch.Tensor, torch.Tensor]:
    """Reverse-time GAE(λ) with torchrl 0.6.0's operation order and dtypes.

    delta_t = r_t + (gamma * not_terminated_t) * V'_t - V_t
    A_t     = delta_t + (lmbda*gamma as a python double, then * not_done_t) * A_{t+1}
    target  = A + V
    """
    if terminated is None:
        terminated = done.clone()
    if not (next_state_value.shape == state_value.shape == reward.shape == done.shape ==
            terminated.shape):
        raise RuntimeError("All input tensors (value, reward and done states) must share a unique shape.")
    nd = state_value.dim()
    td = time_dim if time_dim >= 0 else nd + time_dim
    if td != nd - 2:  # the decorator in torchrl moves time to -2 and back
        mv = lambda x: x.transpose(td, nd - 2)
        adv, tgt = generalized_advantage_estimate(gamma, lmbda, mv(state_value), mv(next_state_value),
                                                  mv(reward), mv(done), mv(terminated), -2)
        return mv(adv), mv(tgt)
    dtype = next_state_value.dtype
    not_done = (~done).int()
    not_terminated = (~terminated).int()
    *batch, steps, feat = not_done.shape
    advantage = torch.empty(*batch, steps, feat, dtype=dtype)
    g_not_terminated = gamma * not_terminated
    delta = reward + (g_not_terminated * next_state_value) - state_value
    discount = lmbda * gamma * not_done
    prev = 0
    for t in reversed(range(steps)):
        prev = advantage[..., t, :] = delta[..., t, :] + (prev * discount[..., t, :])
    value_target = advantage + state_value
    return advantage, value_target


def calculate_advantages(reward: torch.Tensor,
                         state_value: torch.Tensor,
                         next_state_value: torch.Tensor,
                         terminated: torch.Tensor,
                         gamma: float,
                         lmbda: float,
                         normalize_rewards: bool = False,
                         normalize_advantage: bool = False,
                         advantage_scaler: float = 1.0) -> Tuple[torch.Tensor, torch.Tensor]:
    """`PPO.calculate_advantages` — src/entities/algorithms/ppo.py:62-91.

    reward/state_value/next_state_value are [N,T,1]; terminated is [N,T] bool.
    Returns (advantage, value_target), both [N,T,1].
    """
    with torch.no_grad():
        r = reward
        if normalize_rewards:  # ppo.py:66-69 (per env over time, unbiased std, no epsilon)
            r = r - r.mean(dim=1).unsqueeze(1)
            r = (r / r.std(dim=1).unsqueeze(1)) * advantage_scaler
        term = terminated.unsqueeze(-1)  # ppo.py:70
        done = term.clone()
        done[:, -1, :] = True  # ppo.py:72: the rollout always ends at the last step
        adv, tgt = generalized_advantage_estimate(gamma, lmbda, state_value, next_state_value, r, done, term)
        if normalize_advantage:  # ppo.py:81-88
            adv = adv - adv.mean(dim=1).unsqueeze(1)
            adv = (adv / adv.std(dim=1).unsqueeze(1)) * advantage_scaler
            tgt = tgt - tgt.mean(dim=1).unsqueeze(1)
            tgt = (tgt / tgt.std(dim=1).unsqueeze(1)) * advantage_scaler
    return adv, tgt


# ----------------------------------------------------------------------------------
# Networks — src/models/network_block_creator.py:24-86, linear/actor.py:7-33, critic.py:6-25
# Attribute names are kept so that state_dict keys equal the reference's.
# ----------------------------------------------------------------------------------
_ACTS = {"tanh": nn.Tanh, "relu": nn.ReLU}


class MLPBlock(nn.Module):
    """hidden × (Linear → act) → Linear [→ Tanh]; orthogonal init (network_block_creator.py:44-72)."""

    def __init__(self, in_dim: int, hidden: Sequence[int], out_dim: int, activation: str, final_tanh: bool,
                 last_layer_std: float = 0.01):
        super().__init__()
        mods: List[nn.Module] = []
        d = in_dim
        for h in hidden:
            lin = nn.Linear(d, h, bias=True)
            with torch.no_grad():
                nn.init.orthogonal_(lin.weight, math.sqrt(2.0))  # layer_init, :18-21
                lin.bias.fill_(0)  # :51-52
            mods += [lin, _ACTS[activation]()]
            d = h
        self.first_layers = nn.Sequential(*mods)
        self.last_layer = nn.Linear(d, out_dim, bias=True)
        with torch.no_grad():
            nn.init.orthogonal_(self.last_layer.weight, last_layer_std)  # :63-65 (bias left at default init)
        self.last_layer_activation = nn.Tanh() if final_tanh else None

    def forward(self, x):
        y = self.last_layer(self.first_layers(x))
        return self.last_layer_activation(y) if self.last_layer_activation is not None else y


class OracleActor(nn.Module):
    """src/models/linear/actor.py:7-33."""

    def __init__(self, obs_dim: int, act_dim: int, hidden: Sequence[int], activation: str,
                 output_max_value: float = 1.0):
        super().__init__()
        self.actor = MLPBlock(obs_dim, hidden, act_dim, activation, final_tanh=True)
        self.actor_logstd = nn.Parameter(torch.zeros(act_dim))
        self.output_max_value = output_max_value

    def forward(self, x):
        x = x.reshape(len(x), -1)
        mean = self.output_max_value * self.actor(x)
        std = self.actor_logstd.exp()
        return mean, torch.repeat_interleave(std[None, :], x.shape[0], dim=0)


class OracleCritic(nn.Module):
    """src/models/critic.py:6-25 (hidden sizes are an argument; the reference hard-codes [128,128])."""

    def __init__(self, obs_dim: int, hidden: Sequence[int], activation: str):
        super().__init__()
        self.network = MLPBlock(obs_dim, hidden, 1, activation, final_tanh=False)

    def forward(self, x):
        return self.network(x.reshape(len(x), -1))


@dataclass
class OracleConfig:
    obs_dim: int
    act_dim: int
    actor_hidden: Sequence[int]
    critic_hidden: Sequence[int]
    activation: str = "tanh"
    output_max_value: float = 1.0
    learning_rate: float = 1e-4
    batch_size: int = 500
    epochs: int = 10
    gamma: float = 0.99
    lmbda: float = 0.98
    clip_epsilon: float = 0.1
    entropy_eps: float = 1e-4
    max_grad_norm: float = 1.0
    normalize_rewards: bool = False
    normalize_advantage: bool = False
    advantage_scaler: float = 1.0


class OracleAgent:
    """PPOAgent with the MLP nets — src/entities/agents/ppo_agent.py:12-43, agent.py:17-42."""

    def __init__(self, cfg: OracleConfig):
        self.cfg = cfg
        self.networks = nn.ModuleDict()
        self.networks["actor"] = OracleActor(cfg.obs_dim, cfg.act_dim, cfg.actor_hidden, cfg.activation,
                                             cfg.output_max_value)
        self.networks["critic"] = OracleCritic(cfg.obs_dim, cfg.critic_hidden, cfg.activation)
        # ppo_agent.py:15-18 — two independent Adam optimisers; foreach=False = the single-tensor CPU path
        self.optimizers = {
            "actor": torch.optim.Adam(self.networks["actor"].parameters(), lr=cfg.learning_rate, foreach=False),
            "critic": torch.optim.Adam(self.networks["critic"].parameters(), lr=cfg.learning_rate, foreach=False),
        }

    def act(self, state, test_phase: bool = False):
        mean, std = self.networks["actor"](state)
        dist = torch.distributions.Normal(mean, std)
        action = mean if test_phase else dist.sample()
        return action, dist

    def get_state_value(self, state):
        return self.networks["critic"](state)


def ppo_train(agent: OracleAgent, mem: Dict[str, torch.Tensor], perms: Sequence[torch.Tensor],
              max_minibatches: Optional[int] = None) -> List[Tuple[float, float]]:
    """`PPO.train` — src/entities/algorithms/ppo.py:93-154, with the permutations supplied.

    `mem` holds the flattened rollout (flat index n*T + t, ppo.py:99): current_state [M,...],
    action [M,A], action_log_prob [M], advantage [M,1], current_state_value_target [M,1].
    `perms[e]` replaces `torch.randperm(len(memory))` of epoch e (ppo.py:103).
    Returns [(actor_loss, critic_loss)] per minibatch in execution order.
    """
    cfg = agent.cfg
    total = mem["action"].shape[0]
    bsz = cfg.batch_size
    nb = int(total / bsz)  # ppo.py:97-98: the tail beyond nb*bsz is dropped
    losses: List[Tuple[float, float]] = []
    done_mb = 0
    for idx in perms:
        shuffled = {k: v[idx] for k, v in mem.items()}  # ppo.py:104
        for i in range(nb):
            if max_minibatches is not None and done_mb >= max_minibatches:
                return losses
            sl = slice(i * bsz, (i + 1) * bsz)
            obs, act = shuffled["current_state"][sl], shuffled["action"][sl]
            mean, std = agent.networks["actor"](obs)  # ppo.py:110
            dist = torch.distributions.Normal(mean, std)
            new_logp = dist.log_prob(act).sum(dim=1)  # :113
            value = agent.get_state_value(obs)  # :115
            critic_loss = huber_loss(value, shuffled["current_state_value_target"][sl], reduction="mean")  # :117
            agent.optimizers["critic"].zero_grad()
            critic_loss.backward()
            agent.optimizers["critic"].step()  # :120-122
            entropy = dist.entropy().mean()  # :125
            ratio = (new_logp - shuffled["action_log_prob"][sl]).exp()[:, None]  # :126
            adv = shuffled["advantage"][sl]
            s1 = ratio * adv
            s2 = torch.clamp(ratio, 1.0 - cfg.clip_epsilon, 1.0 + cfg.clip_epsilon) * adv
            actor_loss = -torch.min(s1, s2).mean() - entropy * cfg.entropy_eps  # :131-132
            agent.optimizers["actor"].zero_grad()
            actor_loss.backward()
            agent.optimizers["actor"].step()  # :133-135
            # :136-137 clip_grad_norm_ AFTER both steps: scales .grad only, the next zero_grad discards it.
            torch.nn.utils.clip_grad_norm_(agent.networks.parameters(), cfg.max_grad_norm)
            losses.append((actor_loss.detach().item(), critic_loss.detach().item()))
            done_mb += 1
    return losses


def minibatch_grads(agent: OracleAgent, obs, act, old_logp, adv, tgt):
    """Losses and gradients of one minibatch without stepping (autograd of ppo.py:110-132)."""
    cfg = agent.cfg
    for p in agent.networks.parameters():
        p.grad = None
    mean, std = agent.networks["actor"](obs)
    dist = torch.distributions.Normal(mean, std)
    new_logp = dist.log_prob(act).sum(dim=1)
    value = agent.get_state_value(obs)
    critic_loss = huber_loss(value, tgt, reduction="mean")
    critic_loss.backward()
    entropy = dist.entropy().mean()
    ratio = (new_logp - old_logp).exp()[:, None]
    s1 = ratio * adv
    s2 = torch.clamp(ratio, 1.0 - cfg.clip_epsilon, 1.0 + cfg.clip_epsilon) * adv
    actor_loss = -torch.min(s1, s2).mean() - entropy * cfg.entropy_eps
    actor_loss.backward()
    grads = {n: p.grad.detach().clone() for n, p in agent.networks.named_parameters()}
    return actor_loss.item(), critic_loss.item(), grads, new_logp.detach(), value.detach()


# ----------------------------------------------------------------------------------
# Synthetic trajectories — SURVEY.md §8(d)
# ----------------------------------------------------------------------------------
def synthetic_rollout(n_envs: int, steps: int, obs_dim: int, act_dim: int, seed: int,
                      p_term: float = 0.01, reward_f64: bool = False) -> Dict[str, torch.Tensor]:
    """Seeded CPU tensors in the reference's buffer layout (env-major, time-minor; ppo.py:30-50,60)."""
    g = torch.Generator().manual_seed(seed)
    obs = torch.randn(n_envs, steps, 1, obs_dim, generator=g)
    action = torch.randn(n_envs, steps, act_dim, generator=g).clamp_(-3, 3)
    value = torch.randn(n_envs, steps, 1, generator=g)
    terminated = torch.rand(n_envs, steps, generator=g) < p_term
    boot = torch.randn(n_envs, steps, 1, generator=g)
    next_value = torch.cat([value[:, 1:], boot[:, -1:]], dim=1)
    next_value = torch.where(terminated.unsqueeze(-1), boot, next_value)
    reward = torch.randn(n_envs, steps, 1, generator=g, dtype=torch.float64 if reward_f64 else torch.float32)
    return {
        "current_state": obs,
        "current_state_value": value,
        "next_state_value": next_value,
        "action": action,
        "reward": reward,
        "terminated": terminated,
        "truncated": torch.zeros_like(terminated),
    }


# ----------------------------------------------------------------------------------
# Observation normalisation — src/environments/humanoid/running_gym_sequential_vectorized.py:61-92
# ----------------------------------------------------------------------------------
def normalize_state(observation: torch.Tensor, bounds=(0, 22, 45, 175, 253, 270), normalize: bool = True) -> torch.Tensor:
    """observation [N, obs_dim, window] (float64 from gym) -> [N, window, obs_dim] float32."""
    state = observation.clone()
    if normalize:
        edges = [b for b in bounds if b < state.shape[1]] + [state.shape[1]]
        for b, e in zip(edges[:-1], edges[1:]):
            seg = state[:, b:e]
            seg = seg - seg.mean(dim=1).unsqueeze(1)      # :62
            std = seg.std(dim=1).unsqueeze(1)             # :63 (unbiased)
            std[std == 0] = 1                             # :64
            state[:, b:e] = seg / std                     # :65
    return state.to(torch.float32).permute(0, 2, 1)       # :89-91


def rollout(agent: OracleAgent, helper, steps: int, noise: torch.Tensor) -> Dict[str, torch.Tensor]:
    """`PPO.rollout` — src/entities/algorithms/ppo.py:13-60 with the Normal draws supplied (`noise` [T, N, A]:
    action = mean + std * noise is what `Normal.sample()` computes)."""
    helper.reset()
    helper.reset_environment(test_phase=False)
    next_state = helper.get_state(test_phase=False)
    items = {k: [] for k in ("current_state", "current_state_value", "next_state_value", "action", "action_log_prob",
                             "reward", "terminated", "truncated")}
    with torch.no_grad():
        for t in range(steps):
            current_state = torch.clone(next_state)
            value = agent.get_state_value(current_state)
            mean, std = agent.networks["actor"](current_state)
            action = mean + std * noise[t]
            logp = torch.distributions.Normal(mean, std).log_prob(action).sum(dim=1)
            helper.step(action)
            next_state = helper.get_state(test_phase=False)
            next_value = agent.get_state_value(next_state)
            items["current_state"].append(current_state.unsqueeze(1))
            items["current_state_value"].append(value.unsqueeze(1))
            items["next_state_value"].append(next_value.unsqueeze(1))
            items["action"].append(action.unsqueeze(1))
            items["action_log_prob"].append(logp.unsqueeze(1))
            items["reward"].append(torch.tensor(helper.timestep.reward)[:, None].unsqueeze(1))
            items["terminated"].append(torch.tensor(helper.timestep.terminated[:, None]))
            items["truncated"].append(torch.tensor(helper.timestep.truncated[:, None]))
    return {k: torch.cat(v, dim=1) for k, v in items.items()}
