"""Functional entry points of the hot path: GAE, gather, Adam.  Thin wrappers over the C ABI."""
from __future__ import annotations

from typing import Optional, Tuple

import torch

from . import _lib

SHAPE_ERR = "All input tensors (value, reward and done states) must share a unique shape."


def _as_bool_bytes(t: torch.Tensor, name: str) -> torch.Tensor:
    _lib.require_cuda(t, name)
    if t.dtype != torch.bool:
        raise RuntimeError(f"{name} must be a bool tensor, got {t.dtype}")
    return t.contiguous().view(torch.uint8)


def _gae_rows(reward, value, next_value, terminated, done, gamma, lmbda, normalize_rewards=False,
              normalize_advantage=False, advantage_scaler=1.0):
    """All arguments [rows, T] contiguous CUDA tensors (done may be None)."""
    lib = _lib.load()
    rows, steps = value.shape
    adv = torch.empty_like(value)
    tgt = torch.empty_like(value)
    if value.numel() == 0:
        return adv, tgt
    _lib.check(
        lib.b200ppo_gae(_lib.ptr(reward), int(reward.dtype == torch.float64), _lib.ptr(value), _lib.ptr(next_value),
                        _lib.ptr(terminated), _lib.ptr(done), rows, steps, float(gamma), float(lmbda),
                        int(bool(normalize_rewards)), int(bool(normalize_advantage)), float(advantage_scaler),
                        _lib.ptr(adv), _lib.ptr(tgt), _lib.stream_ptr()), "b200ppo_gae")
    return adv, tgt


def generalized_advantage_estimate(gamma: float,
                                   lmbda: float,
                                   state_value: torch.Tensor,
                                   next_state_value: torch.Tensor,
                                   reward: torch.Tensor,
                                   done: torch.Tensor,
                                   terminated: Optional[torch.Tensor] = None,
                                   time_dim: int = -2) -> Tuple[torch.Tensor, torch.Tensor]:
    """Drop-in for torchrl 0.6.0 `generalized_advantage_estimate` (call site
    src/entities/algorithms/ppo.py:76-80): tensors shaped [*B, T, F], returns (advantage, value_target)."""
    if terminated is None:
        terminated = done.clone()
    if not (next_state_value.shape == state_value.shape == reward.shape == done.shape == terminated.shape):
        raise RuntimeError(SHAPE_ERR)
    _lib.require_cuda(state_value, "state_value", torch.float32)
    _lib.require_cuda(next_state_value, "next_state_value", torch.float32)
    _lib.require_cuda(reward, "reward")
    if reward.dtype not in (torch.float32, torch.float64):
        raise RuntimeError(f"reward must be float32 or float64, got {reward.dtype}")
    nd = state_value.dim()
    if nd < 2:
        raise RuntimeError("expected tensors shaped [*B, T, F]")
    td = time_dim if time_dim >= 0 else nd + time_dim
    # kernel layout: [rows, T] with time contiguous; [N, T, 1] (the reference's case) is already that.
    def to_rows(x):
        x = x.movedim(td, -1)  # [..., F, T] when td == nd-2
        return x.contiguous().reshape(-1, x.shape[-1]), x.shape
    v, shp = to_rows(state_value)
    nv, _ = to_rows(next_state_value)
    r, _ = to_rows(reward)
    d, _ = to_rows(_as_bool_bytes(done, "done"))
    t, _ = to_rows(_as_bool_bytes(terminated, "terminated"))
    adv, tgt = _gae_rows(r, v, nv, t, d, gamma, lmbda)
    back = lambda x: x.reshape(shp).movedim(-1, td).contiguous()
    return back(adv), back(tgt)


def calculate_advantages(reward, state_value, next_state_value, terminated, gamma, lmbda, normalize_rewards=False,
                         normalize_advantage=False, advantage_scaler=1.0):
    """Fused `PPO.calculate_advantages` (ppo.py:62-91): reward/value/next_value [N,T,1], terminated [N,T] bool.
    The forced `done[:, -1] = True`, optional reward normalisation, the GAE scan and the optional per-env
    normalisation of both outputs run in ONE kernel.  Returns (advantage, value_target) shaped [N,T,1]."""
    _lib.require_cuda(state_value, "current_state_value", torch.float32)
    _lib.require_cuda(next_state_value, "next_state_value", torch.float32)
    _lib.require_cuda(reward, "reward")
    if not (state_value.shape == next_state_value.shape == reward.shape):
        raise RuntimeError(SHAPE_ERR)
    n, steps = state_value.shape[0], state_value.shape[1]
    if state_value.numel() != n * steps or terminated.numel() != n * steps:
        raise RuntimeError(SHAPE_ERR)
    if reward.dtype not in (torch.float32, torch.float64):
        raise RuntimeError(f"reward must be float32 or float64, got {reward.dtype}")
    term = _as_bool_bytes(terminated, "terminated").reshape(n, steps)
    adv, tgt = _gae_rows(reward.contiguous().reshape(n, steps), state_value.contiguous().reshape(n, steps),
                         next_state_value.contiguous().reshape(n, steps), term, None, gamma, lmbda, normalize_rewards,
                         normalize_advantage, advantage_scaler)
    return adv.reshape(state_value.shape), tgt.reshape(state_value.shape)


def gather_rows(src: torch.Tensor, idx: torch.Tensor) -> torch.Tensor:
    """`src[idx]` along dim 0 for any dtype (bit-exact row copies)."""
    _lib.require_cuda(src, "src")
    _lib.require_cuda(idx, "idx", torch.int64)
    lib = _lib.load()
    src = src.contiguous()
    idx = idx.contiguous()
    n_rows = src.shape[0]
    row_bytes = (src.numel() // max(n_rows, 1)) * src.element_size() if n_rows else 0
    out = torch.empty((idx.numel(), *src.shape[1:]), dtype=src.dtype, device=src.device)
    if idx.numel() == 0 or row_bytes == 0:
        return out
    err = torch.zeros(1, dtype=torch.int32, device=src.device)
    _lib.check(
        lib.b200ppo_gather_rows(_lib.ptr(src), row_bytes, n_rows, _lib.ptr(idx), idx.numel(), _lib.ptr(out),
                                _lib.ptr(err), _lib.stream_ptr()), "b200ppo_gather_rows")
    if int(err.item()) != 0:
        raise IndexError("index out of range in gather_rows")
    return out


def gather_minibatch(idx, obs, action, logp, advantage, target, check: bool = True):
    """The five leaves of `memory[idx]` the update reads (ppo.py:104-124), one launch."""
    lib = _lib.load()
    for name, t in (("obs", obs), ("action", action), ("logp", logp), ("advantage", advantage), ("target", target)):
        _lib.require_cuda(t, name, torch.float32)
    _lib.require_cuda(idx, "idx", torch.int64)
    m = obs.shape[0]
    obs2, act2 = obs.contiguous().reshape(m, -1), action.contiguous().reshape(m, -1)
    cnt = idx.numel()
    o = torch.empty((cnt, obs2.shape[1]), dtype=torch.float32, device=obs.device)
    a = torch.empty((cnt, act2.shape[1]), dtype=torch.float32, device=obs.device)
    lp, ad, tg = (torch.empty(cnt, dtype=torch.float32, device=obs.device) for _ in range(3))
    if cnt == 0:
        return o, a, lp, ad, tg
    err = torch.zeros(1, dtype=torch.int32, device=obs.device)
    _lib.check(
        lib.b200ppo_gather_minibatch(_lib.ptr(idx.contiguous()), cnt, m, _lib.ptr(obs2), obs2.shape[1], _lib.ptr(act2),
                                     act2.shape[1], _lib.ptr(logp.contiguous().reshape(-1)),
                                     _lib.ptr(advantage.contiguous().reshape(-1)),
                                     _lib.ptr(target.contiguous().reshape(-1)), _lib.ptr(o), _lib.ptr(a), _lib.ptr(lp),
                                     _lib.ptr(ad), _lib.ptr(tg), _lib.ptr(err), _lib.stream_ptr()),
        "b200ppo_gather_minibatch")
    if check and int(err.item()) != 0:
        raise IndexError("index out of range in gather_minibatch")
    return o, a, lp, ad, tg


def adam_step_(param, grad, exp_avg, exp_avg_sq, step: int, lr: float, beta1=0.9, beta2=0.999, eps=1e-8):
    """In-place `torch.optim.Adam` single-tensor update (ppo.py:122,135) on flat contiguous fp32 CUDA tensors."""
    lib = _lib.load()
    for name, t in (("param", param), ("grad", grad), ("exp_avg", exp_avg), ("exp_avg_sq", exp_avg_sq)):
        _lib.require_cuda(t, name, torch.float32)
        if not t.is_contiguous():
            raise RuntimeError(f"{name} must be contiguous")
    _lib.check(
        lib.b200ppo_adam_step(_lib.ptr(param), _lib.ptr(grad), 1, param.numel(), _lib.ptr(exp_avg), _lib.ptr(exp_avg_sq),
                              param.numel(), float(lr), float(beta1), float(beta2), float(eps), int(step),
                              _lib.stream_ptr()), "b200ppo_adam_step")
    return param


def polyak_update_(target: torch.Tensor, source: torch.Tensor, tau: float) -> torch.Tensor:
    """In-place `target = target * (1 - tau) + source * tau` on flat contiguous fp32 CUDA tensors
    (`soft_update`, soft_actor_critic.py:12-14, for every tensor of a network at once)."""
    lib = _lib.load()
    for name, t in (("target", target), ("source", source)):
        _lib.require_cuda(t, name, torch.float32)
        if not t.is_contiguous():
            raise RuntimeError(f"{name} must be contiguous")
    if target.numel() != source.numel():
        raise RuntimeError("target and source must have the same number of elements")
    _lib.check(lib.b200ppo_polyak_update(_lib.ptr(target), _lib.ptr(source), target.numel(), float(tau), _lib.stream_ptr()),
               "b200ppo_polyak_update")
    return target


# index ranges of the SymmetricHumanoid observation vector the reference normalises separately
# (src/environments/humanoid/running_gym_sequential_vectorized.py:69-80); the last range ends at obs_dim
HUMANOID_SEGMENTS = (0, 22, 45, 175, 253, 270)


def normalize_state(observation: torch.Tensor, segments=HUMANOID_SEGMENTS, normalize: bool = True) -> torch.Tensor:
    """`EnvironmentHelper.get_state` minus the environment (running_gym_sequential_vectorized.py:61-92): per-env,
    per-frame normalisation of every index range of the observation, cast to fp32, permute to [N, window, obs].
    `observation` is [N, obs_dim, window], float32 or float64 (gym), on the GPU."""
    _lib.require_cuda(observation, "observation")
    if observation.dtype not in (torch.float32, torch.float64) or observation.dim() != 3:
        raise RuntimeError("observation must be a float32/float64 tensor shaped [N, obs_dim, window]")
    lib = _lib.load()
    obs = observation.contiguous()
    n, d, w = obs.shape
    bounds = [int(b) for b in segments if int(b) < d]
    if not bounds or bounds[0] != 0:
        bounds = [0] + bounds
    bounds.append(d)
    seg = torch.tensor(bounds, dtype=torch.int32, device=obs.device)
    out = torch.empty((n, w, d), dtype=torch.float32, device=obs.device)
    _lib.check(
        lib.b200ppo_normalize_obs(_lib.ptr(obs), int(obs.dtype == torch.float64), n, d, w, _lib.ptr(seg), len(bounds) - 1,
                                  int(bool(normalize)), _lib.ptr(out), _lib.stream_ptr()), "b200ppo_normalize_obs")
    return out
