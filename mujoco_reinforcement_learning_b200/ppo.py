"""PPO algorithm object with the reference's methods (src/entities/algorithms/ppo.py:10-159):
`rollout`, `calculate_advantages`, `train`, `_iterate` — the numerics run in `libb200ppo.so`."""
from __future__ import annotations

from typing import List, Optional, Sequence

import torch

from . import functional as F
from .agent import PPOAgent
from .config import Run
from .memory import RolloutMemory


def eng_in_dim(engine) -> int:
    return int(getattr(engine, "obs_dim", -1))


class PPO:

    def __init__(self, environment_helper, agent: PPOAgent, match_reference_rng: bool = False):
        self.environment_helper = environment_helper
        self.agent = agent
        # The reference draws and discards one Normal sample per minibatch (ppo.py:110 via ppo_agent.py:40) from
        # the same CPU generator as `torch.randperm`.  True replays those draws so that, from the same seed,
        # epoch k's permutation equals the reference's; False (default) skips the wasted work.
        self.match_reference_rng = match_reference_rng
        self.last_losses: Optional[torch.Tensor] = None
        self.last_episode_losses = (float("nan"), float("nan"))

    @property
    def run(self) -> Run:
        return self.environment_helper.run

    # ------------------------------------------------------------------------------------------------
    @torch.no_grad()
    def rollout(self, noise: Optional[torch.Tensor] = None) -> RolloutMemory:
        """ppo.py:13-60 with pre-allocated [N, T, ...] device buffers written in place (no per-step
        containers, no final concat) and one fused inference call per step for V(s), pi(s), log-prob.
        `noise` ([T, N, act_dim] standard-normal draws) replaces the device RNG, for reproducible comparisons."""
        helper, run = self.environment_helper, self.run
        helper.reset()
        helper.reset_environment(test_phase=False)
        next_state = helper.get_state(test_phase=False)
        n, steps = len(next_state), run.environment_config.maximum_timesteps
        dev = torch.device(run.device)
        next_state = next_state.to(dev, torch.float32)
        buf = {
            "current_state": torch.empty((n, steps, *next_state.shape[1:]), dtype=torch.float32, device=dev),
            "current_state_value": torch.empty((n, steps, 1), dtype=torch.float32, device=dev),
            "next_state_value": torch.empty((n, steps, 1), dtype=torch.float32, device=dev),
            "action": torch.empty((n, steps, self.agent.engine.act_dim), dtype=torch.float32, device=dev),
            "action_log_prob": torch.empty((n, steps), dtype=torch.float32, device=dev),
            "reward": torch.empty((n, steps, 1), dtype=torch.float64, device=dev),
            "terminated": torch.empty((n, steps), dtype=torch.bool, device=dev),
            "truncated": torch.empty((n, steps), dtype=torch.bool, device=dev),
        }
        eng = self.agent.engine
        # the native step writes s_t, V(s_t), a_t, log pi(a_t) and V(s') of step t - 1 straight into the buffers; it takes the
        # state as the flat [N, window * obs] row the networks see (network_block_creator.py flattens the same way)
        flat_state = next_state.reshape(n, -1).shape[1] == eng_in_dim(self.agent.engine)
        for t in range(steps):
            current_state = next_state
            if flat_state:
                eps = torch.randn((n, eng.act_dim), dtype=torch.float32, device=dev) if noise is None else noise[t].to(dev)
                action = eng.rollout_step(current_state.reshape(n, -1), eps, t, buf)
            else:
                action, logp, value = self.agent.act_fused(current_state, None if noise is None else noise[t].to(dev))
            helper.step(action)
            next_state = helper.get_state(test_phase=False).to(dev, torch.float32)
            if not flat_state:
                buf["current_state"][:, t] = current_state
                buf["current_state_value"][:, t] = value
                buf["action"][:, t] = action
                buf["action_log_prob"][:, t] = logp
                buf["next_state_value"][:, t] = self.agent.get_state_value(next_state)
            ts = helper.timestep
            buf["reward"][:, t, 0] = torch.as_tensor(ts.reward, dtype=torch.float64).to(dev)
            buf["terminated"][:, t] = torch.as_tensor(ts.terminated, dtype=torch.bool).to(dev)
            buf["truncated"][:, t] = torch.as_tensor(ts.truncated, dtype=torch.bool).to(dev)
        if flat_state:
            eng.rollout_step(next_state.reshape(n, -1), None, steps, buf)  # V of the final state closes next_state_value
        return RolloutMemory(buf, (n, steps))

    # ------------------------------------------------------------------------------------------------
    @torch.no_grad()
    def calculate_advantages(self, memory):
        """ppo.py:62-91 — one fused kernel; writes `current_state_value_target` and `advantage`."""
        run = self.run
        adv, tgt = F.calculate_advantages(memory["reward"], memory["current_state_value"], memory["next_state_value"],
                                          memory["terminated"], run.ppo_config.gamma, run.ppo_config.lmbda,
                                          normalize_rewards=run.normalize_rewards,
                                          normalize_advantage=run.ppo_config.normalize_advantage,
                                          advantage_scaler=run.ppo_config.advantage_scaler)
        memory["current_state_value_target"] = tgt
        memory["advantage"] = adv

    # ------------------------------------------------------------------------------------------------
    def draw_permutations(self, total: int, epochs: int, batches_per_epoch: int, batch_size: int) -> torch.Tensor:
        """`torch.randperm(len(memory))` per epoch on the CPU default generator (ppo.py:103)."""
        perms: List[torch.Tensor] = []
        a = self.agent.engine.act_dim
        for _ in range(epochs):
            perms.append(torch.randperm(total))
            if self.match_reference_rng:
                z, o = torch.zeros(batch_size, a), torch.ones(batch_size, a)
                for _ in range(batches_per_epoch):
                    torch.normal(z, o)  # the discarded dist.sample() of ppo.py:110
        return torch.stack(perms)

    def train(self, memory, perms: Optional[torch.Tensor] = None, max_minibatches_per_epoch: int = 0):
        """ppo.py:93-154.  `perms` ([epochs, N*T] int64) overrides the permutations drawn here."""
        run = self.run
        batch_size = int(run.training_config.batch_size)
        epochs = int(run.training_config.epochs_per_iteration)
        total = int(run.environment_config.maximum_timesteps * run.environment_config.num_envs)
        batches_per_epoch = int(total / batch_size)  # ppo.py:97-98
        flat = memory.view(-1)  # flat index n*T + t (ppo.py:99)
        if len(flat) != total:
            raise RuntimeError(f"memory holds {len(flat)} samples, config says {total}")
        eng = self.agent.engine
        if perms is None:
            perms = self.draw_permutations(total, epochs, batches_per_epoch, batch_size)
        perms = perms.to(eng.device, non_blocking=True)
        hp = eng.hparams(self.agent.optimizers["actor"].param_groups[0]["lr"],
                         self.agent.optimizers["critic"].param_groups[0]["lr"], run.ppo_config.clip_epsilon,
                         run.ppo_config.entropy_eps, self.agent.optimizers["actor"].param_groups[0]["betas"],
                         self.agent.optimizers["actor"].param_groups[0]["eps"])
        losses = eng.train(flat["current_state"], flat["action"], flat["action_log_prob"], flat["advantage"],
                           flat["current_state_value_target"], perms, batch_size, hp, max_minibatches_per_epoch)
        # ppo.py:136-137 clips gradients only after both optimiser steps: nothing to reproduce.
        self.last_losses = losses
        if losses.numel():
            m = losses.view(epochs, -1, 2).mean(dim=1).mean(dim=0).tolist()  # the one host sync of the call
            self.last_episode_losses = (m[0], m[1])
        if run.dynamic_config.current_episode < 2500:  # ppo.py:146-148
            for scheduler in self.agent.schedulers.values():
                scheduler.step()
        return self.last_episode_losses

    def _iterate(self):
        memory = self.rollout()
        self.calculate_advantages(memory)
        self.train(memory)
