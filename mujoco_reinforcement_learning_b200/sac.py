"""Soft Actor-Critic update step on the CUDA library.

Mirrors `SoftActorCritic.train` (src/entities/algorithms/soft_actor_critic.py:33-118), `soft_update` / `hard_update`
(:12-19), `QNetwork` (src/models/linear/q_network.py:7-38) and `SoftActorCriticAgent`
(src/entities/agents/soft_actor_critic_agent.py:10-38, with the MLP actor / Q network instead of the Transformer
variants it binds as committed).  The arithmetic that dominates — the MLP forwards and backwards of the actor and of
the twin Q networks (including dQ/da through the Q networks for the reparameterised policy gradient), the replay
gather, Adam and the Polyak average — runs in the library's kernels through the same C ABI as the PPO path
(`b200ppo_mlp_forward/backward`, `b200ppo_gather_rows`, `b200ppo_adam_step`, `b200ppo_polyak_update`); the per-sample
scalar glue (log-prob, min, TD target, MSE) is a handful of element-wise device ops recorded by autograd.
"""
from __future__ import annotations

import math
import os
from typing import Dict, Optional

import torch
from torch import nn
from torch.nn import ModuleDict

from .config import Run
from .functional import adam_step_, gather_rows, polyak_update_
from .models import Actor, ActorCriticEngine, Critic, create_network


class QNetwork(nn.Module):
    """Twin MLP Q functions on cat[state.flatten(1), action] — src/models/linear/q_network.py:7-38."""

    def __init__(self, run: Optional[Run] = None):
        super().__init__()
        run = run or Run.instance()
        nc = run.network_config
        config = {"final_activation": None, "activation": nc.activation_class, "hidden_layer_count": nc.num_linear_layers,
                  "shapes": nc.linear_hidden_shapes}
        in_dim = int(nc.input_shape * 2 + nc.output_shape)  # q_network.py:19-20 (a window of two frames)
        self.first_network = create_network(config, input_shape=in_dim, output_shape=1, normalize_at_the_end=False,
                                            use_bias=nc.use_bias, use_batchnorm=nc.use_batch_norm)
        self.second_network = create_network(config, input_shape=in_dim, output_shape=1, normalize_at_the_end=False,
                                             use_bias=nc.use_bias, use_batchnorm=nc.use_batch_norm)

    def forward(self, state: torch.Tensor, action: torch.Tensor):
        x = torch.cat([state.reshape(len(state), -1), action], 1)
        return self.first_network(x), self.second_network(x)


class _AsActor(nn.Module):
    """Presents a plain NetworkBlock in the engine's 'actor' slot (its log-std slot stays an untrained zero)."""

    def __init__(self, block):
        super().__init__()
        self.actor = block
        self.actor_logstd = nn.Parameter(torch.zeros(1), requires_grad=False)
        self.output_max_value, self.output_shape = 1.0, 1


class _AsCritic(nn.Module):
    def __init__(self, block):
        super().__init__()
        self.network = block


class _FlatAdam:
    """One Adam launch over a contiguous range of an engine's flat buffers (torch single-tensor semantics per element).
    Gradients accumulate straight into a flat buffer: every parameter's `.grad` is a view of it."""

    def __init__(self, engine: ActorCriticEngine, begin: int, end: int, lr: float):
        self.engine, self.begin, self.end, self.lr = engine, begin, end, lr
        self.grad = torch.zeros(engine.n_params, dtype=torch.float32, device=engine.device)
        self.step_count = 0
        self.params = [(p, off) for p, off in engine.slots if begin <= off < end and p.requires_grad]

    def zero_grad(self):
        self.grad[self.begin:self.end].zero_()
        for p, off in self.params:
            p.grad = self.grad[off:off + p.numel()].view(p.shape)

    def clip_grad_norm_(self, max_norm: float) -> torch.Tensor:
        """torch.nn.utils.clip_grad_norm_ (L2): grads *= min(1, max_norm / (total_norm + 1e-6))."""
        g = self.grad[self.begin:self.end]
        total = torch.linalg.vector_norm(g)
        g.mul_(torch.clamp(max_norm / (total + 1e-6), max=1.0))
        return total

    def step(self):
        self.step_count += 1
        eng, b, e = self.engine, self.begin, self.end
        adam_step_(eng.flat[b:e], self.grad[b:e], eng.exp_avg[b:e], eng.exp_avg_sq[b:e], self.step_count, self.lr)

    # torch.optim.Adam's state_dict layout (state[i] = {step, exp_avg, exp_avg_sq}, one param group), so that the files
    # `Agent.save` writes (agent.py:51-55) load into stock torch optimisers and back
    def state_dict(self):
        state = {}
        if self.step_count > 0:
            for i, (p, off) in enumerate(self.params):
                sl = slice(off, off + p.numel())
                state[i] = {"step": torch.tensor(float(self.step_count)), "exp_avg": self.engine.exp_avg[sl].view(p.shape).clone(),
                            "exp_avg_sq": self.engine.exp_avg_sq[sl].view(p.shape).clone()}
        group = {"lr": self.lr, "betas": (0.9, 0.999), "eps": 1e-8, "weight_decay": 0, "amsgrad": False, "maximize": False,
                 "foreach": False, "capturable": False, "differentiable": False, "fused": None, "params": list(range(len(self.params)))}
        return {"state": state, "param_groups": [group]}

    def load_state_dict(self, sd):
        self.lr = float(sd["param_groups"][0]["lr"])
        self.step_count = 0
        for i, (p, off) in enumerate(self.params):
            st = sd["state"].get(i)
            if st is None:
                continue
            sl = slice(off, off + p.numel())
            self.engine.exp_avg[sl].copy_(st["exp_avg"].reshape(-1).to(self.engine.device))
            self.engine.exp_avg_sq[sl].copy_(st["exp_avg_sq"].reshape(-1).to(self.engine.device))
            self.step_count = int(float(st["step"]))


class _ExponentialLR:
    """`ExponentialLR(optimizer, gamma)` for a _FlatAdam: `step()` multiplies its learning rate (the reference steps one per
    optimiser after every trained iteration, soft_actor_critic.py:199-201)."""

    def __init__(self, optimizer: _FlatAdam, gamma: float):
        self.optimizer, self.gamma = optimizer, gamma

    def step(self):
        self.optimizer.lr *= self.gamma


class SoftActorCriticAgent:
    """networks: actor, online_critic, target_critic; optimizers: actor, online_critic (soft_actor_critic_agent.py:12-35)."""

    def __init__(self, run: Optional[Run] = None, max_batch: Optional[int] = None):
        self.run = run = run or Run.instance()
        if run.sac_config.automatic_entropy_tuning:
            raise NotImplementedError("automatic entropy tuning is off on the reference's default path and not implemented")
        if run.gemm_precision != "fp32":
            raise NotImplementedError("the SAC step runs on the fp32 kernels")
        nc = run.network_config
        if run.environment_config.window_length != 2:
            raise RuntimeError("the reference's linear QNetwork reads input_shape * 2 state features: window_length must be 2")
        mb = max_batch or int(run.training_config.batch_size)
        self.networks = ModuleDict()
        self.networks["actor"] = Actor(run)            # construction order = order of random draws of the reference
        self.networks["online_critic"] = QNetwork(run)
        self.networks["target_critic"] = QNetwork(run)
        # engines: the actor shares its context with an unused value head; each Q pair fills one context
        self._value_stub = Critic(Run(network_config=type(nc)(input_shape=nc.input_shape, output_shape=nc.output_shape,
                                                               activation_class=nc.activation_class, num_linear_layers=1,
                                                               linear_hidden_shapes=[8]),
                                      environment_config=run.environment_config, device=run.device))
        Run._instance = run  # the stub's Run() must not replace the caller's singleton
        self.engine_pi = ActorCriticEngine(self.networks["actor"], self._value_stub, max_batch=mb, device=run.device)
        oc, tc = self.networks["online_critic"], self.networks["target_critic"]
        self._wrap = [_AsActor(oc.first_network), _AsCritic(oc.second_network), _AsActor(tc.first_network), _AsCritic(tc.second_network)]
        self.engine_q = ActorCriticEngine(self._wrap[0], self._wrap[1], max_batch=mb, device=run.device)
        self.engine_qt = ActorCriticEngine(self._wrap[2], self._wrap[3], max_batch=mb, device=run.device)
        lr = run.training_config.learning_rate
        self.optimizers: Dict[str, _FlatAdam] = {
            "actor": _FlatAdam(self.engine_pi, 0, self.engine_pi.n_actor, lr),
            "online_critic": _FlatAdam(self.engine_q, 0, self.engine_q.n_params, lr),
        }
        self.schedulers = {name: _ExponentialLR(opt, 0.999) for name, opt in self.optimizers.items()}  # soft_actor_critic_agent.py:30-34
        for p in self.networks["target_critic"].parameters():
            p.requires_grad_(False)

    # -- checkpoints (agent.py:47-72): same files, same state_dict keys as the reference's Agent.save / load ------------
    def save(self):
        run = self.run
        ep = run.dynamic_config.current_episode
        os.makedirs(f"{run.experiment_path}/networks/{ep}", exist_ok=True)
        torch.save(self.networks.state_dict(), f"{run.experiment_path}/networks/{ep}/networks.pth")
        for name, opt in self.optimizers.items():
            torch.save(opt.state_dict(), f"{run.experiment_path}/networks/{ep}/optimizer_{name}.pth")

    def load(self):
        run = self.run
        load_path = f"{run.experiment_path}/networks/{run.dynamic_config.current_episode}"
        if not os.path.exists(load_path):
            raise ValueError("the current path not exist!")
        self.networks.load_state_dict(torch.load(f"{load_path}/networks.pth", map_location=run.device))
        for eng in (self.engine_pi, self.engine_q, self.engine_qt):
            eng.bind_views()
        for name, opt in self.optimizers.items():
            opt.load_state_dict(torch.load(f"{load_path}/optimizer_{name}.pth", map_location=run.device))

    def act(self, state: torch.Tensor, return_dist: bool = False, test_phase: bool = False, noise: Optional[torch.Tensor] = None):
        """agent.py:26-42 with `rsample()`: action = mean + std * eps."""
        means, stds = self.networks["actor"](state)
        if test_phase:
            action = means
        else:
            eps = noise if noise is not None else torch.randn_like(means)
            action = means + eps * stds
        if return_dist:
            return action, (means, stds)
        return action


def hard_update(agent: SoftActorCriticAgent):
    """soft_actor_critic.py:17-19 on the flat buffers."""
    agent.engine_qt.ensure_bound(); agent.engine_q.ensure_bound()
    agent.engine_qt.flat.copy_(agent.engine_q.flat)


def _log_prob_sum(mean, std, value):
    var = std ** 2
    return (-((value - mean) ** 2) / (2 * var) - std.log() - math.log(math.sqrt(2 * math.pi))).sum(dim=1)[:, None]


class SoftActorCritic:
    """`train(memory, update_count)` — soft_actor_critic.py:33-118 (one minibatch per call, entropy tuning off)."""

    def __init__(self, environment_helper, agent: SoftActorCriticAgent):
        self.environment_helper, self.agent = environment_helper, agent
        hard_update(agent)  # :30
        self.alpha = float(environment_helper.run.sac_config.alpha)

    def train(self, memory, update_count: int, idx: Optional[torch.Tensor] = None, noise: Optional[torch.Tensor] = None):
        """memory: mapping of `[N, T, ...]` CUDA leaves `current_state`, `next_state`, `reward`, `action`, `is_alive`.
        idx: the permutation of :39 (default: `torch.randperm` on the CPU generator, as the reference draws it);
        noise [2, B, A]: the standard-normal draws of the two `rsample()` calls (default: `torch.randn` on the device)."""
        run: Run = self.environment_helper.run
        ag = self.agent
        B = int(run.training_config.batch_size)
        n_lead = memory["reward"].dim() - 1
        flat = {k: memory[k].reshape(-1, *memory[k].shape[n_lead:]) for k in ("current_state", "next_state", "reward", "action", "is_alive")}
        M = flat["reward"].shape[0]
        dev = flat["reward"].device
        if idx is None:
            idx = torch.randperm(M)
        idx = idx.to(dev)[:B].contiguous()
        # :40-44 — only the first batch_size rows of memory[idx] are read; the reward statistics cover the whole memory
        r_all = flat["reward"]
        r_mean = r_all.mean()
        r_std = (r_all - r_mean).std()
        s, s2, a = (gather_rows(flat[k].contiguous(), idx) for k in ("current_state", "next_state", "action"))
        r = (gather_rows(r_all.contiguous(), idx) - r_mean) / r_std
        mask = gather_rows(flat["is_alive"].contiguous().to(torch.uint8), idx).to(torch.float32)
        A = a.shape[1]
        if noise is None:
            noise = torch.randn((2, B, A), dtype=torch.float32, device=dev)
        actor, online, target = ag.networks["actor"], ag.networks["online_critic"], ag.networks["target_critic"]
        gamma = float(run.sac_config.gamma)
        with torch.no_grad():  # :50-59
            mean2, std2 = actor(s2)
            a2 = mean2 + noise[0] * std2
            logp2 = _log_prob_sum(mean2, std2, a2)
            q1t, q2t = target(s2, a2)
            y = (r + mask * gamma * (torch.min(q1t, q2t) - self.alpha * logp2)).to(torch.float32)
        opt_q, opt_pi = ag.optimizers["online_critic"], ag.optimizers["actor"]
        q1, q2 = online(s, a)  # :60-68
        qf1_loss, qf2_loss = torch.nn.functional.mse_loss(q1, y), torch.nn.functional.mse_loss(q2, y)
        opt_q.zero_grad()
        (qf1_loss + qf2_loss).backward()
        opt_q.clip_grad_norm_(float(run.ppo_config.max_grad_norm))  # :73-74
        opt_q.step()
        mean, std = actor(s)  # :77-86
        a_pi = mean + noise[1] * std
        q1p, q2p = online(s, a_pi)
        logp = _log_prob_sum(mean, std, a_pi)
        min_q_pi = torch.min(q1p, q2p)
        policy_loss = (self.alpha * logp - min_q_pi).mean()
        opt_pi.zero_grad()
        policy_loss.backward()
        opt_pi.clip_grad_norm_(float(run.ppo_config.max_grad_norm))  # :90-91
        opt_pi.step()
        if update_count % int(run.sac_config.target_update_interval) == 0:  # :109-111, every tensor of both Q networks at once
            polyak_update_(ag.engine_qt.flat, ag.engine_q.flat, float(run.sac_config.tau))
        return torch.stack([qf1_loss.detach(), qf2_loss.detach(), policy_loss.detach(), min_q_pi.mean().detach(),
                            torch.zeros((), device=dev)])
