"""§8f rows next to the hot path: the in-place rollout buffer writer (PPO.rollout, ppo.py:13-60) and the observation
normalisation feeding the policy (running_gym_sequential_vectorized.py:61-92), against the oracle / reference goldens."""
import numpy as np
import pytest
import torch

import mujoco_reinforcement_learning_b200 as pkg
from oracle import ppo_oracle as O
from tests._fake_env import FakeHelper
from tests._util import RTOL_FP32, assert_close, load_golden
from tests.test_update_gpu import make_pair

pytestmark = pytest.mark.gpu
DEV = "cuda"


def test_normalize_state_vs_reference_golden():
    g = load_golden("obsnorm")
    obs = torch.from_numpy(g["in_observation"])
    out = pkg.normalize_state(obs.to(DEV))
    assert out.shape == g["state"].shape and out.dtype == torch.float32
    assert_close(out, g["state"], RTOL_FP32, "normalized state (float64 input)")
    out32 = pkg.normalize_state(obs.float().to(DEV))
    assert_close(out32, g["state"], 1e-4, "normalized state (float32 input)")  # statistics in fp32 instead of fp64


@pytest.mark.parametrize("n,d,w,f64", [(1, 376, 5, True), (64, 376, 1, True), (7, 348, 5, False), (3, 30, 2, True), (4096, 376, 5, True)])
def test_normalize_state_vs_oracle(n, d, w, f64):
    g = torch.Generator().manual_seed(n + d + w)
    obs = torch.randn(n, d, w, generator=g, dtype=torch.float64 if f64 else torch.float32) * 2 + 1
    ref = O.normalize_state(obs)
    out = pkg.normalize_state(obs.to(DEV))
    assert_close(out, ref, RTOL_FP32 if f64 else 1e-4, "normalize_state")
    plain = pkg.normalize_state(obs.to(DEV), normalize=False)
    assert torch.equal(plain.cpu(), obs.float().permute(0, 2, 1))


def test_rollout_writes_the_reference_buffer():
    N, T, D, A = 12, 40, 10, 3
    oracle, agent, run = make_pair(D, A, [32, 32], [32, 32], "tanh", n_envs=N, steps=T, max_batch=128)
    noise = torch.randn(T, N, A, generator=torch.Generator().manual_seed(5))
    ref = O.rollout(oracle, FakeHelper(run, N, D, A, seed=3), T, noise)
    algo = pkg.PPO(FakeHelper(run, N, D, A, seed=3), agent)
    mem = algo.rollout(noise=noise)
    assert mem.batch_size == (N, T)
    assert mem["reward"].dtype == torch.float64 and mem["terminated"].dtype == torch.bool
    for k in ("current_state", "current_state_value", "next_state_value", "action", "action_log_prob", "reward"):
        assert tuple(mem[k].shape) == tuple(ref[k].shape), k
        assert_close(mem[k], ref[k], 1e-4, k)  # 40 steps of closed-loop feedback through fp32 policies
    assert torch.equal(mem["terminated"].cpu(), ref["terminated"])
    # and the buffer feeds the advantage pipeline + update directly
    algo.calculate_advantages(mem)
    adv, tgt = O.calculate_advantages(ref["reward"], ref["current_state_value"], ref["next_state_value"], ref["terminated"],
                                      0.99, 0.98)
    assert_close(mem["advantage"], adv, 1e-4, "advantage from the rollout buffer")
    run.training_config.batch_size = 96
    run.training_config.epochs_per_iteration = 1
    algo.train(mem)
    assert np.isfinite(algo.last_episode_losses).all()


def test_rollout_through_the_one_launch_step_bf16():
    """bf16 context, 256-wide nets: every environment step is b200ppo_rollout_step's tensor-core launch writing slice [:, t]
    of the buffers; V(s') of step t is the very value written as V(s) of step t + 1 (ppo.py:21,27-29)."""
    from tests._util import RTOL_BF16, rel_l2
    N, T, D, A = 300, 6, 40, 4
    oracle, agent, run = make_pair(D, A, [256, 256], [256, 256], "tanh", n_envs=N, steps=T, max_batch=512, precision="bf16")
    noise = torch.randn(T, N, A, generator=torch.Generator().manual_seed(5))
    ref = O.rollout(oracle, FakeHelper(run, N, D, A, seed=3), T, noise)
    algo = pkg.PPO(FakeHelper(run, N, D, A, seed=3), agent)
    mem = algo.rollout(noise=noise)
    assert torch.equal(mem["next_state_value"][:, :-1], mem["current_state_value"][:, 1:])
    for k in ("current_state", "current_state_value", "next_state_value", "action", "action_log_prob", "reward"):
        assert tuple(mem[k].shape) == tuple(ref[k].shape), k
        e = rel_l2(mem[k], ref[k])
        assert e <= 1.5 * RTOL_BF16, (k, e)  # six steps of closed-loop feedback through bf16 policies
    assert torch.equal(mem["terminated"].cpu(), ref["terminated"])
