"""Independent float64 numpy restatement — TEST INFRASTRUCTURE ONLY.

Explicit loops and hand-derived gradients; shares no code with `ppo_oracle.py`, so an
error in one shows up as a disagreement.  Formula sources:
 * GAE: torchrl 0.6.0 `generalized_advantage_estimate` (call site
   src/entities/algorithms/ppo.py:76-80) — delta/discount/recurrence/target.
 * log-prob / entropy: torch `distributions/normal.py` (`log_prob`, `entropy`), used at
   src/entities/algorithms/ppo.py:113,125.
 * losses: src/entities/algorithms/ppo.py:117-119,126-132.
 * Adam: torch `optim/adam.py::_single_tensor_adam` (non-capturable, no amsgrad, no
   weight decay), constructed at src/entities/agents/ppo_agent.py:15-18.
"""
from __future__ import annotations

import math
from typing import Dict, List, Sequence, Tuple

import numpy as np


def gae_naive(reward, value, next_value, done, terminated, gamma: float, lmbda: float):
    """reward/value/next_value: [N,T] float; done/terminated: [N,T] bool.  Returns (adv, target) float64."""
    reward = np.asarray(reward, dtype=np.float64)
    value = np.asarray(value, dtype=np.float64)
    next_value = np.asarray(next_value, dtype=np.float64)
    n_envs, steps = reward.shape
    adv = np.zeros((n_envs, steps), dtype=np.float64)
    for n in range(n_envs):
        carry = 0.0
        for t in range(steps - 1, -1, -1):
            nt = 0.0 if terminated[n, t] else 1.0
            nd = 0.0 if done[n, t] else 1.0
            delta = reward[n, t] + gamma * nt * next_value[n, t] - value[n, t]
            carry = delta + lmbda * gamma * nd * carry
            adv[n, t] = carry
    return adv, adv + value


def normalize_rows_naive(x, scaler: float):
    """(x - mean_T) / std_T(unbiased) * scaler per env — ppo.py:66-69 / :81-88."""
    x = np.asarray(x, dtype=np.float64)
    out = np.empty_like(x)
    for n in range(x.shape[0]):
        m = x[n].sum() / x.shape[1]
        c = x[n] - m
        var = (c * c).sum() / (x.shape[1] - 1)
        out[n] = c / math.sqrt(var) * scaler
    return out


def _act(z, kind):
    return np.tanh(z) if kind == "tanh" else np.maximum(z, 0.0)


def _dact_from_out(h, kind):
    return 1.0 - h * h if kind == "tanh" else (h > 0).astype(np.float64)


def mlp_forward(x, weights: Sequence[Tuple[np.ndarray, np.ndarray]], kind: str):
    """weights = [(W[out,in], b[out]), ...]; hidden layers use `kind`, last layer is linear.
    Returns (pre-activation of last layer, list of layer inputs)."""
    acts = [x]
    h = x
    for W, b in weights[:-1]:
        h = _act(h @ W.T + b, kind)
        acts.append(h)
    W, b = weights[-1]
    return h @ W.T + b, acts


def mlp_backward(dz_last, acts, weights, kind):
    """Back-propagate dL/dz_last through the block; returns [(dW, db)] in layer order."""
    grads: List[Tuple[np.ndarray, np.ndarray]] = [None] * len(weights)
    dz = dz_last
    for li in range(len(weights) - 1, -1, -1):
        grads[li] = (dz.T @ acts[li], dz.sum(axis=0))
        if li > 0:
            dh = dz @ weights[li][0]
            dz = dh * _dact_from_out(acts[li], kind)
    return grads


def ppo_minibatch_naive(actor_w, logstd, critic_w, obs, act, old_logp, adv, tgt, kind: str,
                        clip_eps: float, ent_coef: float, out_max: float = 1.0):
    """One minibatch of ppo.py:109-135 in float64: losses, new log-prob, value and all gradients."""
    f = lambda a: np.asarray(a, dtype=np.float64)
    actor_w = [(f(W), f(b)) for W, b in actor_w]
    critic_w = [(f(W), f(b)) for W, b in critic_w]
    logstd, obs, act = f(logstd), f(obs), f(act)
    old_logp, adv, tgt = f(old_logp).reshape(-1), f(adv).reshape(-1), f(tgt).reshape(-1)
    B, A = act.shape
    # actor forward
    z3, a_acts = mlp_forward(obs, actor_w, kind)
    th = np.tanh(z3)
    mean = out_max * th
    sigma = np.exp(logstd)
    diff = act - mean
    logp = (-(diff ** 2) / (2.0 * sigma ** 2) - logstd - math.log(math.sqrt(2.0 * math.pi))).sum(axis=1)
    entropy = (0.5 + 0.5 * math.log(2.0 * math.pi) + logstd).sum() / A  # mean over [B,A] of a row-constant
    ratio = np.exp(logp - old_logp)
    s1 = ratio * adv
    s2 = np.clip(ratio, 1.0 - clip_eps, 1.0 + clip_eps) * adv
    actor_loss = -np.minimum(s1, s2).mean() - ent_coef * entropy
    # d(-mean(min(s1,s2)))/d ratio, reproducing autograd's tie rule (min → half/half) and the
    # inclusive clamp mask (d clamp/dx = 1 for lo <= x <= hi).
    in_range = ((ratio >= 1.0 - clip_eps) & (ratio <= 1.0 + clip_eps)).astype(np.float64)
    w1 = np.where(s1 < s2, 1.0, np.where(s1 > s2, 0.0, 0.5))
    g_ratio = -(w1 * adv + (1.0 - w1) * adv * in_range) / B
    g_logp = g_ratio * ratio
    d_mean = g_logp[:, None] * diff / sigma ** 2
    d_logstd = (g_logp[:, None] * (diff ** 2 / sigma ** 2 - 1.0)).sum(axis=0) - ent_coef / A
    dz3 = d_mean * out_max * (1.0 - th * th)
    actor_g = mlp_backward(dz3, a_acts, actor_w, kind)
    # critic
    v, c_acts = mlp_forward(obs, critic_w, kind)
    v = v.reshape(-1)
    e = v - tgt
    hub = np.where(np.abs(e) < 1.0, 0.5 * e * e, np.abs(e) - 0.5)
    critic_loss = hub.mean()
    dv = np.clip(e, -1.0, 1.0) / B
    critic_g = mlp_backward(dv[:, None], c_acts, critic_w, kind)
    return {
        "actor_loss": actor_loss, "critic_loss": critic_loss, "logp": logp, "value": v, "mean": mean,
        "actor_grads": actor_g, "logstd_grad": d_logstd, "critic_grads": critic_g,
    }


def adam_step_naive(p, g, m, v, step: int, lr: float, b1=0.9, b2=0.999, eps=1e-8):
    """One `_single_tensor_adam` update (float64); `step` is the 1-based step count after increment."""
    m = b1 * m + (1.0 - b1) * g
    v = b2 * v + (1.0 - b2) * g * g
    bc1 = 1.0 - b1 ** step
    bc2 = 1.0 - b2 ** step
    denom = np.sqrt(v) / math.sqrt(bc2) + eps
    p = p - (lr / bc1) * m / denom
    return p, m, v
