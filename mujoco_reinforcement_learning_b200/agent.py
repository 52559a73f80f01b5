"""PPO agent with the reference's surface: `networks`, `optimizers`, `schedulers`, `act`, `get_state_value`,
`save` / `load` (src/entities/agents/agent.py:14-72, src/entities/agents/ppo_agent.py:10-43)."""
from __future__ import annotations

from os import makedirs, path
from typing import Dict, Optional

import torch
from torch.nn import ModuleDict
from torch.optim.lr_scheduler import ExponentialLR

from . import _lib
from .config import Run
from .functional import adam_step_
from .models import Actor, ActorCriticEngine, Critic


class FusedAdam(torch.optim.Adam):
    """`torch.optim.Adam` whose state lives in the engine's flat buffers and whose `step()` runs the CUDA kernel.

    `state_dict()` / `load_state_dict()` keep torch's layout (`state[i] = {step, exp_avg, exp_avg_sq}`,
    `param_groups`), so checkpoints move between this implementation and the reference (agent.py:51-55,69-71).
    """

    def __init__(self, engine: ActorCriticEngine, net_id: int, lr: float):
        self.engine, self.net_id = engine, net_id
        params = [p for p, _ in engine.module_slots(net_id)]  # torch's numbering: module.parameters() order
        super().__init__(params, lr=lr, foreach=False)
        self.bind_state()

    def bind_state(self, copy_from_state: bool = False):
        eng = self.engine
        for p, off in eng.named_slots(self.net_id):
            st = self.state[p]
            for key, flat in (("exp_avg", eng.exp_avg), ("exp_avg_sq", eng.exp_avg_sq)):
                view = flat[off:off + p.numel()].view(p.shape)
                if copy_from_state and key in st and st[key].data_ptr() != view.data_ptr():
                    view.copy_(st[key].to(view.device))
                st[key] = view
            if copy_from_state and "step" in st:
                eng.adam_steps[self.net_id] = int(float(st["step"]))
            st["step"] = torch.tensor(float(eng.adam_steps[self.net_id]))

    def sync_step_from_engine(self):
        """The per-parameter `step` entries torch's state_dict carries all mirror the engine's count for this optimiser."""
        for p, _ in self.engine.named_slots(self.net_id):
            self.state[p]["step"] = torch.tensor(float(self.engine.adam_steps[self.net_id]))

    def state_dict(self):
        self.sync_step_from_engine()
        return super().state_dict()

    def load_state_dict(self, state_dict):
        super().load_state_dict(state_dict)
        self.bind_state(copy_from_state=True)

    @torch.no_grad()
    def step(self, closure=None):
        """Per-tensor path for user-driven loops (loss.backward(); optimizer.step()).  The fused trainer
        (`PPO.train`) does not come through here."""
        group = self.param_groups[0]
        b1, b2 = group["betas"]
        t = self.engine.adam_steps[self.net_id] + 1  # one count per optimiser, kept in the engine (also what `train` continues from)
        stepped = False
        for p, _ in self.engine.named_slots(self.net_id):
            if p.grad is None:
                continue
            st = self.state[p]
            adam_step_(p.data.view(-1), p.grad.contiguous().view(-1), st["exp_avg"].view(-1), st["exp_avg_sq"].view(-1), t,
                       group["lr"], b1, b2, group["eps"])
            st["step"] = torch.tensor(float(t))
            stepped = True
        if stepped:
            self.engine.adam_steps[self.net_id] = t



class NormalLike:
    """The slice of `torch.distributions.Normal` the reference touches (`log_prob`, `entropy`, `sample`)."""

    def __init__(self, mean, std):
        self._d = torch.distributions.Normal(mean, std)
        self.mean, self.stddev = mean, std

    def log_prob(self, value):
        return self._d.log_prob(value)

    def entropy(self):
        return self._d.entropy()

    def sample(self):
        return self._d.sample()


class PPOAgent:

    def __init__(self, run: Optional[Run] = None, max_batch: Optional[int] = None):
        self.run = run or Run.instance()
        if self.run is None:
            raise RuntimeError("construct a Run configuration first")
        self.networks: ModuleDict = ModuleDict()
        self.optimizers: Dict[str, FusedAdam] = dict()
        self.schedulers: Dict[str, ExponentialLR] = dict()
        self.max_batch = max_batch
        self.initialize_networks()

    def initialize_networks(self):
        run = self.run
        self.networks["actor"] = Actor(run)    # ppo_agent.py:13-14 construction order (actor first)
        self.networks["critic"] = Critic(run)
        mb = self.max_batch or max(int(run.training_config.batch_size), int(run.environment_config.num_envs))
        self.engine = ActorCriticEngine(self.networks["actor"], self.networks["critic"], max_batch=mb, device=run.device,
                                        precision=run.gemm_precision)
        lr = run.training_config.learning_rate
        self.optimizers["actor"] = FusedAdam(self.engine, 0, lr)   # ppo_agent.py:15-18
        self.optimizers["critic"] = FusedAdam(self.engine, 1, lr)
        self.schedulers["actor"] = ExponentialLR(self.optimizers["actor"], gamma=0.999)   # ppo_agent.py:21-22
        self.schedulers["critic"] = ExponentialLR(self.optimizers["critic"], gamma=0.999)

    def get_state_value(self, state: torch.Tensor) -> torch.Tensor:
        return self.networks["critic"](state)

    def act(self, state: torch.Tensor, return_dist: bool = False, test_phase: bool = False):
        """ppo_agent.py:27-43: action = mean (test) or a Normal sample; optionally the distribution."""
        means, stds = self.networks["actor"](state)
        dist = NormalLike(means, stds)
        action = means if test_phase else dist.sample()
        if return_dist:
            return action, dist
        return action

    @torch.no_grad()
    def act_fused(self, state: torch.Tensor, noise: Optional[torch.Tensor] = None, test_phase: bool = False):
        """Rollout fast path (ppo.py:22-26 in one native call): returns (action, log_prob, value)."""
        if not test_phase and noise is None:
            noise = torch.randn((len(state), self.engine.act_dim), dtype=torch.float32, device=state.device)
        action, logp, value, _ = self.engine.policy_infer(state, None if test_phase else noise)
        return action, logp, value

    @torch.no_grad()
    def evaluate(self, state: torch.Tensor, actions: torch.Tensor):
        """(new log-prob [B], entropy, value [B,1]) — ppo.py:109-115,125."""
        return self.engine.evaluate(state, actions)

    # -- checkpoints (agent.py:47-72) ---------------------------------------------------------------------
    def save(self):
        run = self.run
        ep = run.dynamic_config.current_episode
        makedirs(f"{run.experiment_path}/networks/{ep}", exist_ok=True)
        torch.save(self.networks.state_dict(), f"{run.experiment_path}/networks/{ep}/networks.pth")
        for name, opt in self.optimizers.items():
            torch.save(opt.state_dict(), f"{run.experiment_path}/networks/{ep}/optimizer_{name}.pth")

    def load(self):
        run = self.run
        ep = run.dynamic_config.current_episode
        load_path = f"{run.experiment_path}/networks/{ep}"
        if not path.exists(load_path):
            load_path = f"{run.experiment_path}/networks/best_results/{ep}"
        if not path.exists(load_path):
            raise ValueError("the current iteration does not exist")
        self.networks.load_state_dict(torch.load(f"{load_path}/networks.pth"))
        self.engine.ensure_bound()
        for name, opt in self.optimizers.items():
            opt.load_state_dict(torch.load(f"{load_path}/optimizer_{name}.pth"))
