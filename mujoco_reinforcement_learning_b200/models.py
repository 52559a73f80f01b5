"""Actor / Critic modules with the reference's call shapes and `state_dict` keys, backed by the CUDA library.

Mirrors `src/models/network_block_creator.py:24-102`, `src/models/linear/actor.py:7-33` and
`src/models/critic.py:6-25`.  Every parameter is a view into ONE flat fp32 CUDA buffer laid out as
`include/b200ppo.h` describes, so the fused update step (forward, losses, backward, Adam) runs on the
whole parameter set with no per-tensor work, while `state_dict()` / `load_state_dict()` keep the
reference's key names (`actor.actor.first_layers.0.weight`, ..., `critic.network.last_layer.bias`).
"""
from __future__ import annotations

import ctypes as C
import functools
import math
from typing import List, Optional, Sequence

import torch
from torch import nn

from . import _lib
from .config import Run

_ACT_CODES = {nn.Tanh: _lib.ACT_TANH, nn.ReLU: _lib.ACT_RELU}


def layer_init(layer: nn.Linear, std: float = math.sqrt(2.0)) -> nn.Linear:
    """network_block_creator.py:18-21 — orthogonal weight, bias untouched."""
    torch.nn.init.orthogonal_(layer.weight, std)
    return layer


class NetworkBlock(nn.Module):
    """hidden_layer_count x (Linear -> activation) -> Linear [-> Tanh].

    Same constructor order of random draws as the reference (`Linear` default init, then `orthogonal_`),
    so the same torch seed yields the same initial weights.  Batch-norm, skip connections and end
    normalisation are not on the PPO path and are rejected.
    """

    def __init__(self, config: dict, input_shape: int, output_shape: int, normalize_at_the_end: bool = False,
                 use_bias: bool = True, use_batchnorm: bool = False, last_layer_std: float = 0.01):
        super().__init__()
        if normalize_at_the_end or use_batchnorm or config.get("skip_connection", False) or not use_bias:
            raise NotImplementedError("only the plain biased MLP of the PPO path is implemented")
        act = config["activation"]
        if act not in _ACT_CODES:
            raise NotImplementedError(f"activation {act} not supported (Tanh / ReLU)")
        final = config.get("final_activation")
        if final not in (None, nn.Tanh):
            raise NotImplementedError("final activation must be None or Tanh")
        self.hidden_layer_count = int(config["hidden_layer_count"])
        self.activation_code = _ACT_CODES[act]
        self.final_tanh = final is nn.Tanh
        self.in_dim = int(input_shape)
        self.dims: List[int] = [int(s) for s in config["shapes"][:self.hidden_layer_count]] + [int(output_shape)]
        layers: List[nn.Module] = []
        d = self.in_dim
        for h in self.dims[:-1]:
            lin = nn.Linear(d, h, bias=True)
            with torch.no_grad():
                layer_init(lin)
                lin.bias.fill_(0)
            layers += [lin, act()]
            d = h
        self.first_layers = nn.Sequential(*layers)
        self.last_layer = nn.Linear(d, self.dims[-1], bias=True)
        with torch.no_grad():
            layer_init(self.last_layer, std=last_layer_std)
        self.last_layer_activation = nn.Tanh() if self.final_tanh else None
        self._engine: Optional["ActorCriticEngine"] = None
        self._net_id = -1

    def linears(self) -> List[nn.Linear]:
        return [m for m in self.first_layers if isinstance(m, nn.Linear)] + [self.last_layer]

    def desc(self, out_scale: float = 1.0) -> _lib.MlpDesc:
        d = _lib.MlpDesc()
        d.n_layers = len(self.dims)
        d.in_dim = self.in_dim
        for i, w in enumerate(self.dims):
            d.dims[i] = w
        d.activation = self.activation_code
        d.final_tanh = int(self.final_tanh)
        d.out_scale = float(out_scale)
        return d

    def forward(self, x_in: torch.Tensor) -> torch.Tensor:
        if self._engine is None:
            raise RuntimeError("NetworkBlock is not bound to an ActorCriticEngine (construct it through PPOAgent)")
        return self._engine.mlp(self._net_id, x_in)


def create_network(config, input_shape, output_shape, normalize_at_the_end: bool = False, use_bias: bool = True,
                   use_batchnorm: bool = False, last_layer_std: float = 0.01) -> NetworkBlock:
    """network_block_creator.py:89-102."""
    return NetworkBlock(config, input_shape, output_shape, normalize_at_the_end, use_bias, use_batchnorm,
                        last_layer_std)


class Actor(nn.Module):
    """MLP Gaussian policy — src/models/linear/actor.py:7-33.  `forward(x) -> (mean[B,A], std[B,A])`."""

    def __init__(self, run: Optional[Run] = None):
        super().__init__()
        run = run or Run.instance()
        nc = run.network_config
        config = {"final_activation": nn.Tanh, "activation": nc.activation_class,
                  "hidden_layer_count": nc.num_linear_layers, "shapes": nc.linear_hidden_shapes}
        # last_layer_std is NOT forwarded: the reference's Actor never passes it, so its last layer always gets the
        # create_network default of 0.01 (linear/actor.py:17-23)
        self.actor = create_network(config, int(nc.input_shape * run.environment_config.window_length), nc.output_shape,
                                    False, nc.use_bias, nc.use_batch_norm)
        self.actor_logstd = nn.Parameter(torch.zeros(nc.output_shape))
        self.output_max_value = float(nc.output_max_value)
        self.output_shape = int(nc.output_shape)

    def forward(self, x: torch.Tensor):
        x = x.reshape(len(x), -1)
        mean = self.actor(x)  # output_max_value * tanh(.) is fused into the last GEMM's epilogue
        std = self.actor_logstd[:self.output_shape].exp()
        return mean, torch.repeat_interleave(std[None, :], x.shape[0], dim=0)

    def act(self, x: torch.Tensor):
        return self.forward(x)


class Critic(nn.Module):
    """MLP value network — src/models/critic.py:6-25.  Hidden sizes default to the reference's hard-coded [128, 128]
    (so its `networks.pth` / `optimizer_critic.pth` load at default config); `critic_hidden_shapes` overrides them.  One
    stated deviation: the input is flattened like the actor's, so a [B, W, obs] window works (the reference's critic takes
    `input_shape` without the window factor and broadcasts)."""

    def __init__(self, run: Optional[Run] = None):
        super().__init__()
        run = run or Run.instance()
        nc = run.network_config
        hidden = list(nc.critic_hidden_shapes if nc.critic_hidden_shapes is not None else [128, 128])
        config = {"final_activation": None, "activation": nc.activation_class, "hidden_layer_count": len(hidden),
                  "shapes": hidden}
        self.network = create_network(config, input_shape=int(nc.input_shape * run.environment_config.window_length),
                                      output_shape=1, normalize_at_the_end=False, use_bias=True)

    def forward(self, x: torch.Tensor):
        return self.network(x.reshape(len(x), -1))


class _MLPFunction(torch.autograd.Function):
    """autograd bridge: forward / backward of one NetworkBlock through the C ABI."""

    @staticmethod
    def forward(ctx, engine, net_id, x, *params):
        out, saved = engine._mlp_forward_raw(net_id, x, need_saved=True)
        ctx.engine, ctx.net_id = engine, net_id
        ctx.save_for_backward(x, out, saved)
        ctx.needs_x = x.requires_grad
        return out

    @staticmethod
    def backward(ctx, grad_out):
        x, out, saved = ctx.saved_tensors
        engine, net_id = ctx.engine, ctx.net_id
        gparams, gx = engine._mlp_backward_raw(net_id, x, out, saved, grad_out.contiguous(), ctx.needs_x)
        return (None, None, gx, *gparams)


def _on_engine_device(method):
    """Run an engine entry point with the engine's GPU current: kernels launch on that device's current stream even when
    the caller's current device is another one (a context on cuda:1 driven from cuda:0)."""
    @functools.wraps(method)
    def wrapper(self, *args, **kwargs):
        with torch.cuda.device(self.device):
            return method(self, *args, **kwargs)
    return wrapper


class ActorCriticEngine:
    """Owns the native context, the flat parameter buffer and the flat Adam state of one actor-critic pair."""

    def __init__(self, actor: Actor, critic: Critic, max_batch: int, device="cuda", precision: str = "fp32"):
        if not torch.cuda.is_available():
            raise RuntimeError("mujoco_reinforcement_learning_b200 needs a CUDA device: there is no CPU fallback")
        self.lib = _lib.load()
        self.device = torch.device(device)
        self.actor, self.critic = actor, critic
        self.max_batch = int(max_batch)
        self.precision = {"fp32": _lib.PREC_FP32, "bf16": _lib.PREC_BF16}[precision]
        self._ctx = C.c_void_p()
        a_desc, c_desc = actor.actor.desc(actor.output_max_value), critic.network.desc()
        with torch.cuda.device(self.device):
            _lib.check(self.lib.b200ppo_create(C.byref(a_desc), C.byref(c_desc), self.max_batch, self.precision,
                                               C.byref(self._ctx)), "b200ppo_create")
        self.n_params = int(self.lib.b200ppo_param_count(self._ctx))
        self.n_actor = int(self.lib.b200ppo_actor_param_count(self._ctx))
        self.flat = torch.zeros(self.n_params, dtype=torch.float32, device=self.device)
        self.exp_avg = torch.zeros_like(self.flat)
        self.exp_avg_sq = torch.zeros_like(self.flat)
        # Adam step count per optimiser (0 actor, 1 critic): the ONE source of truth for both update paths — the fused
        # trainer (`train`, which steps both optimisers once per minibatch) and `FusedAdam.step()`
        self.adam_steps = [0, 0]
        self.obs_dim = actor.actor.in_dim
        self.act_dim = actor.output_shape
        # (parameter, element offset) in flat-buffer order == nn.Module.parameters() order
        self.slots = []
        for net_id, block in ((0, actor.actor), (1, critic.network)):
            for li, lin in enumerate(block.linears()):
                self.slots.append((lin.weight, int(self.lib.b200ppo_param_offset(self._ctx, net_id, li, 0))))
                self.slots.append((lin.bias, int(self.lib.b200ppo_param_offset(self._ctx, net_id, li, 1))))
            if net_id == 0:
                self.slots.append((actor.actor_logstd,
                                   int(self.lib.b200ppo_param_offset(self._ctx, 0, len(block.dims), 0))))
            block._engine, block._net_id = self, net_id
        self.bind_views(copy_from_modules=True)

    @property
    def adam_step(self) -> int:
        """Common step count of the two optimisers (what the fused trainer continues from)."""
        if self.adam_steps[0] != self.adam_steps[1]:
            raise RuntimeError(f"actor and critic optimisers are at different Adam steps {self.adam_steps}: the fused trainer "
                               "steps both per minibatch and needs them equal")
        return self.adam_steps[0]

    @adam_step.setter
    def adam_step(self, value: int):
        self.adam_steps = [int(value), int(value)]

    def __del__(self):
        try:
            if getattr(self, "_ctx", None) is not None and self._ctx.value:
                self.lib.b200ppo_destroy(self._ctx)
                self._ctx = C.c_void_p()
        except Exception:
            pass

    # -- parameter storage -------------------------------------------------------------------------------
    def bind_views(self, copy_from_modules: bool = False):
        """(Re)point every nn.Parameter at its slot of the flat buffer."""
        with torch.no_grad():
            for p, off in self.slots:
                view = self.flat[off:off + p.numel()].view(p.shape)
                if copy_from_modules or p.data_ptr() != view.data_ptr():
                    view.copy_(p.detach().to(self.device))
                    p.data = view

    def params_are_bound(self) -> bool:
        base = self.flat.data_ptr()
        return all(p.data_ptr() == base + 4 * off for p, off in self.slots)

    def ensure_bound(self):
        if not self.params_are_bound():
            self.bind_views()

    def segment(self, net_id: int):
        return (0, self.n_actor) if net_id == 0 else (self.n_actor, self.n_params)

    def named_slots(self, net_id: int):
        """(parameter, offset) of one net in flat-buffer order (logstd last for the actor)."""
        n_actor_slots = 2 * len(self.actor.actor.dims) + 1
        return self.slots[:n_actor_slots] if net_id == 0 else self.slots[n_actor_slots:]

    def module_slots(self, net_id: int):
        """(parameter, offset) in `module.parameters()` order — the order torch.optim.Adam numbers its state in
        (the actor's own `actor_logstd` comes before its sub-module's tensors)."""
        off = {id(p): o for p, o in self.slots}
        module = self.actor if net_id == 0 else self.critic
        return [(p, off[id(p)]) for p in module.parameters()]

    # -- raw calls ---------------------------------------------------------------------------------------
    def _check_x(self, x: torch.Tensor) -> torch.Tensor:
        _lib.require_cuda(x, "input", torch.float32)
        x = x.reshape(len(x), -1).contiguous()
        if x.shape[1] != self.obs_dim:
            raise RuntimeError(f"expected {self.obs_dim} input features, got {x.shape[1]}")
        if x.shape[0] > self.max_batch:
            raise RuntimeError(f"batch {x.shape[0]} exceeds the engine's max_batch {self.max_batch}")
        return x

    @_on_engine_device
    def _mlp_forward_raw(self, net_id: int, x: torch.Tensor, need_saved: bool):
        self.ensure_bound()
        x = self._check_x(x)
        B = x.shape[0]
        out_dim = self.act_dim if net_id == 0 else 1
        out = torch.empty((B, out_dim), dtype=torch.float32, device=x.device)
        saved = None
        if need_saved:
            saved = torch.empty(max(1, int(self.lib.b200ppo_saved_size(self._ctx, net_id, B))), dtype=torch.float32,
                                device=x.device)
        _lib.check(self.lib.b200ppo_mlp_forward(self._ctx, net_id, _lib.ptr(self.flat), _lib.ptr(x), B, _lib.ptr(out),
                                                _lib.ptr(saved), _lib.stream_ptr()), "b200ppo_mlp_forward")
        return out, saved

    @_on_engine_device
    def _mlp_backward_raw(self, net_id, x, out, saved, grad_out, need_gx: bool):
        x = self._check_x(x)
        B = x.shape[0]
        slots = [s for s in self.named_slots(net_id) if s[0] is not self.actor.actor_logstd]
        beg, _ = self.segment(net_id)
        end = max(off + p.numel() for p, off in slots)
        gseg = torch.empty(end - beg, dtype=torch.float32, device=x.device)
        gx = torch.empty_like(x) if need_gx else None
        _lib.check(self.lib.b200ppo_mlp_backward(self._ctx, net_id, _lib.ptr(self.flat), _lib.ptr(x), _lib.ptr(out),
                                                 _lib.ptr(saved), _lib.ptr(grad_out), B, _lib.ptr(gseg), _lib.ptr(gx),
                                                 _lib.stream_ptr()), "b200ppo_mlp_backward")
        gparams = [gseg[off - beg:off - beg + p.numel()].view(p.shape) for p, off in slots]
        return gparams, gx

    def mlp(self, net_id: int, x: torch.Tensor) -> torch.Tensor:
        block = self.actor.actor if net_id == 0 else self.critic.network
        params = [p for lin in block.linears() for p in (lin.weight, lin.bias)]
        if torch.is_grad_enabled() and (x.requires_grad or any(p.requires_grad for p in params)):
            return _MLPFunction.apply(self, net_id, x, *params)
        return self._mlp_forward_raw(net_id, x, need_saved=False)[0]

    # -- fused paths -------------------------------------------------------------------------------------
    @_on_engine_device
    def policy_infer(self, obs: torch.Tensor, noise: Optional[torch.Tensor], want_value: bool = True):
        """K6: one call for actor + critic forward, sampling and log-prob (ppo.py:22-26)."""
        self.ensure_bound()
        obs = self._check_x(obs)
        B = obs.shape[0]
        mean = torch.empty((B, self.act_dim), dtype=torch.float32, device=obs.device)
        action = torch.empty_like(mean)
        logp = torch.empty(B, dtype=torch.float32, device=obs.device)
        value = torch.empty((B, 1), dtype=torch.float32, device=obs.device) if want_value else None
        if noise is not None:
            _lib.require_cuda(noise, "noise", torch.float32)
            noise = noise.contiguous()
        _lib.check(self.lib.b200ppo_policy_infer(self._ctx, _lib.ptr(self.flat), _lib.ptr(obs), B, _lib.ptr(noise),
                                                 _lib.ptr(mean), _lib.ptr(value), _lib.ptr(action), _lib.ptr(logp),
                                                 _lib.stream_ptr()), "b200ppo_policy_infer")
        return action, logp, value, mean

    @_on_engine_device
    def evaluate(self, obs: torch.Tensor, actions: torch.Tensor):
        """ppo.py:109-115,125: (new log-prob [B], entropy scalar, value [B,1]) without autograd."""
        self.ensure_bound()
        obs = self._check_x(obs)
        B = obs.shape[0]
        actions = _lib.require_cuda(actions, "actions", torch.float32).reshape(B, -1).contiguous()
        logp = torch.empty(B, dtype=torch.float32, device=obs.device)
        value = torch.empty((B, 1), dtype=torch.float32, device=obs.device)
        ent = torch.empty((), dtype=torch.float32, device=obs.device)
        _lib.check(self.lib.b200ppo_evaluate(self._ctx, _lib.ptr(self.flat), _lib.ptr(obs), _lib.ptr(actions), B,
                                             _lib.ptr(logp), _lib.ptr(value), _lib.ptr(ent), _lib.stream_ptr()),
                   "b200ppo_evaluate")
        return logp, ent, value

    def hparams(self, lr_actor, lr_critic, clip_epsilon, entropy_eps, betas=(0.9, 0.999), eps=1e-8) -> _lib.HParams:
        return _lib.HParams(float(lr_actor), float(lr_critic), float(betas[0]), float(betas[1]), float(eps),
                            float(clip_epsilon), float(entropy_eps))

    @_on_engine_device
    def minibatch_grads(self, obs, action, old_logp, advantage, target, hp: _lib.HParams):
        """Losses and flat gradient of one minibatch (no optimiser step)."""
        self.ensure_bound()
        obs = self._check_x(obs)
        B = obs.shape[0]
        f = lambda t, n: _lib.require_cuda(t, n, torch.float32).contiguous()
        action, old_logp = f(action, "action").reshape(B, -1), f(old_logp, "old_logp").reshape(B)
        advantage, target = f(advantage, "advantage").reshape(B), f(target, "target").reshape(B)
        grads = torch.empty(self.n_params, dtype=torch.float32, device=obs.device)
        losses = torch.empty(2, dtype=torch.float32, device=obs.device)
        _lib.check(self.lib.b200ppo_minibatch_grads(self._ctx, _lib.ptr(self.flat), _lib.ptr(obs), _lib.ptr(action),
                                                    _lib.ptr(old_logp), _lib.ptr(advantage), _lib.ptr(target), B,
                                                    C.byref(hp), _lib.ptr(grads), _lib.ptr(losses), _lib.stream_ptr()),
                   "b200ppo_minibatch_grads")
        return losses, grads

    @_on_engine_device
    def rollout_step(self, obs: torch.Tensor, noise: Optional[torch.Tensor], t: int, buf: dict) -> torch.Tensor:
        """One environment step written into the rollout's [N, T, ...] buffers (ppo.py:20-49); returns buf['action'][:, t].
        t == T: only next_state_value[:, T - 1] from the final state."""
        self.ensure_bound()
        obs = self._check_x(obs)
        N, T = buf["action"].shape[0], buf["action"].shape[1]
        if noise is not None:
            noise = _lib.require_cuda(noise, "noise", torch.float32).contiguous()
        for k in ("current_state", "current_state_value", "next_state_value", "action", "action_log_prob"):
            assert buf[k].is_contiguous() and buf[k].dtype == torch.float32 and buf[k].device == obs.device, k
        _lib.check(self.lib.b200ppo_rollout_step(self._ctx, _lib.ptr(self.flat), _lib.ptr(obs), N, _lib.ptr(noise), int(t), int(T),
                                                 _lib.ptr(buf["current_state"]), _lib.ptr(buf["current_state_value"]),
                                                 _lib.ptr(buf["next_state_value"]), _lib.ptr(buf["action"]),
                                                 _lib.ptr(buf["action_log_prob"]), _lib.stream_ptr()), "b200ppo_rollout_step")
        return buf["action"][:, t] if t < T else None

    def set_fp32_terms(self, terms: int) -> None:
        """fp32 contexts: 3 = three bf16 terms per operand value (24-bit operands, the 1e-5 variant, default); 2 = two scaled
        fp16 terms (22-bit operands, faster; include/b200ppo.h)."""
        _lib.check(self.lib.b200ppo_set_fp32_terms(self._ctx, int(terms)), "b200ppo_set_fp32_terms")

    @_on_engine_device
    def debug_activations(self, net: int, kind: int, layer: int, rows: int) -> torch.Tensor:
        """Test hook: bf16 intermediates of the last bf16 minibatch as fp32 (kind 0: H_layer, kind 1: dL/dz_layer)."""
        dims = (self.actor.actor if net == 0 else self.critic.network).dims
        out = torch.empty(rows, dims[layer], dtype=torch.float32, device=self.flat.device)
        _lib.check(self.lib.b200ppo_debug_activations(self._ctx, net, kind, layer, rows, _lib.ptr(out), _lib.stream_ptr()),
                   "b200ppo_debug_activations")
        return out

    def grads_by_name(self, grads: torch.Tensor, named_parameters):
        """Split a flat gradient into {name: tensor} for an iterable of (name, parameter)."""
        off = {id(p): o for p, o in self.slots}
        return {n: grads[off[id(p)]:off[id(p)] + p.numel()].view(p.shape) for n, p in named_parameters}

    @_on_engine_device
    def train(self, obs, action, old_logp, advantage, target, perms, batch: int, hp: _lib.HParams,
              max_minibatches_per_epoch: int = 0, rank_sliced_perms: bool = False, check_errors: bool = True) -> torch.Tensor:
        """`PPO.train` inner loops (ppo.py:101-140) in one native call.  Returns device losses [epochs*nb, 2].
        rank_sliced_perms: `perms` is [epochs, M // batch, batch // world] — only the permutation slots this rank
        consumes (`distributed.slice_perms_for_rank`) instead of the global [epochs, M]."""
        self.ensure_bound()
        f = lambda t, n: _lib.require_cuda(t, n, torch.float32).contiguous()
        if obs is None:  # observations live in the ranks' shared bf16 tables (distributed.share_rollout)
            M = action.shape[0]
        else:
            M = obs.shape[0]
            obs = f(obs, "current_state").reshape(M, -1)
            if obs.shape[1] != self.obs_dim:
                raise RuntimeError(f"expected {self.obs_dim} observation features, got {obs.shape[1]}")
        action, old_logp = f(action, "action").reshape(M, -1), f(old_logp, "action_log_prob").reshape(M)
        advantage, target = f(advantage, "advantage").reshape(M), f(target, "current_state_value_target").reshape(M)
        perms = _lib.require_cuda(perms, "perms", torch.int64).contiguous()
        if rank_sliced_perms:
            per_epoch = (M // int(batch)) * (int(batch) // max(int(getattr(self, "world", 1)), 1))
            if per_epoch == 0 or perms.numel() % per_epoch != 0:
                raise RuntimeError(f"rank-sliced perms must hold a multiple of {per_epoch} indices, got {perms.numel()}")
            perms = perms.reshape(-1, per_epoch)
        else:
            perms = perms.reshape(-1, M)
        epochs = perms.shape[0]
        nb = M // int(batch)
        if max_minibatches_per_epoch > 0:
            nb = min(nb, max_minibatches_per_epoch)
        losses = torch.zeros((epochs * nb, 2), dtype=torch.float32, device=action.device)
        step = C.c_int64(self.adam_step)
        _lib.check(self.lib.b200ppo_set_perm_layout(self._ctx, 1 if rank_sliced_perms else 0), "b200ppo_set_perm_layout")
        try:
            _lib.check(self.lib.b200ppo_train(self._ctx, _lib.ptr(self.flat), _lib.ptr(self.exp_avg), _lib.ptr(self.exp_avg_sq),
                                              C.byref(step), _lib.ptr(obs) if obs is not None else None, _lib.ptr(action), _lib.ptr(old_logp),
                                              _lib.ptr(advantage), _lib.ptr(target), M, _lib.ptr(perms), epochs, int(batch),
                                              int(max_minibatches_per_epoch), C.byref(hp), _lib.ptr(losses),
                                              _lib.stream_ptr()), "b200ppo_train")
        finally:
            if rank_sliced_perms:
                _lib.check(self.lib.b200ppo_set_perm_layout(self._ctx, 0), "b200ppo_set_perm_layout")
        self.adam_step = int(step.value)
        if check_errors:  # deferred device-side errors (bad permutation entry, peer that never arrived): one stream sync per call
            _lib.check(self.lib.b200ppo_poll_error(self._ctx, _lib.stream_ptr()), "b200ppo_train")
        return losses
