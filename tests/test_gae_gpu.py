"""K1 parity: CUDA GAE / calculate_advantages (through the C ABI) vs the oracle and the reference goldens."""
import numpy as np
import pytest
import torch

import mujoco_reinforcement_learning_b200 as pkg
from oracle import ppo_oracle as O
from tests._util import RTOL_FP32, assert_close, load_golden

pytestmark = pytest.mark.gpu
DEV = "cuda"


@pytest.mark.parametrize("name", ["adv_plain", "adv_norm", "adv_rnorm", "adv_f64"])
def test_calculate_advantages_vs_reference_golden(name):
    g = load_golden(name)
    gamma, lmbda, nr, na, sc = g["cfg"]
    c = lambda k: torch.from_numpy(g[k]).to(DEV)
    adv, tgt = pkg.calculate_advantages(c("in_reward"), c("in_current_state_value"), c("in_next_state_value"),
                                        c("in_terminated"), gamma, lmbda, bool(nr), bool(na), sc)
    assert adv.shape == g["advantage"].shape and adv.dtype == torch.float32
    assert_close(adv, g["advantage"], RTOL_FP32, "advantage")
    assert_close(tgt, g["value_target"], RTOL_FP32, "value_target")


def _rand(n, t, f, seed, f64=False):
    g = torch.Generator().manual_seed(seed)
    v, vn = torch.randn(n, t, f, generator=g), torch.randn(n, t, f, generator=g)
    r = torch.randn(n, t, f, generator=g, dtype=torch.float64 if f64 else torch.float32)
    term = torch.rand(n, t, f, generator=g) < 0.05
    done = term | (torch.rand(n, t, f, generator=g) < 0.05)
    return v, vn, r, done, term


@pytest.mark.parametrize("n,t,f", [(1, 1, 1), (3, 7, 1), (5, 128, 1), (4, 131, 1), (2, 1024, 1), (7, 260, 1),
                                   (3, 33, 2), (1, 2048, 1), (129, 64, 1)])
@pytest.mark.parametrize("f64", [False, True])
def test_gae_vs_oracle(n, t, f, f64):
    v, vn, r, done, term = _rand(n, t, f, 100 + n + t, f64)
    a_ref, t_ref = O.generalized_advantage_estimate(0.99, 0.98, v, vn, r, done, term)
    a, tg = pkg.generalized_advantage_estimate(0.99, 0.98, v.to(DEV), vn.to(DEV), r.to(DEV), done.to(DEV), term.to(DEV))
    assert a.shape == a_ref.shape and a.dtype == torch.float32
    assert_close(a, a_ref, RTOL_FP32, "advantage")
    assert_close(tg, t_ref, RTOL_FP32, "target")


def test_gae_terminated_defaults_to_done_and_time_dim():
    v, vn, r, done, _ = _rand(4, 40, 1, 7)
    a_ref, t_ref = O.generalized_advantage_estimate(0.9, 0.8, v, vn, r, done)
    a, tg = pkg.generalized_advantage_estimate(0.9, 0.8, v.to(DEV), vn.to(DEV), r.to(DEV), done.to(DEV))
    assert_close(a, a_ref)
    assert_close(tg, t_ref)
    tr = lambda x: x.transpose(0, 1).contiguous()
    a2, t2 = pkg.generalized_advantage_estimate(0.9, 0.8, tr(v).to(DEV), tr(vn).to(DEV), tr(r).to(DEV),
                                                tr(done).to(DEV), time_dim=0)
    assert_close(a2, tr(a_ref))
    assert_close(t2, tr(t_ref))


def test_gae_known_answers():
    f = lambda x: torch.tensor(x, dtype=torch.float32, device=DEV).reshape(1, -1, 1)
    b = lambda x: torch.tensor(x, dtype=torch.bool, device=DEV).reshape(1, -1, 1)
    r, v, vn = [1.0, -2.0, 3.0], [0.3, 0.2, 0.1], [9.0, 9.0, 9.0]
    a, _ = pkg.generalized_advantage_estimate(0.99, 0.98, f(v), f(vn), f(r), b([True] * 3), b([True] * 3))
    np.testing.assert_allclose(a.cpu().reshape(-1).numpy(), np.array(r) - np.array(v), atol=1e-6)
    a, _ = pkg.generalized_advantage_estimate(1.0, 1.0, f(v), f(vn), f(r), b([False] * 3), b([False] * 3))
    d = np.array(r) + np.array(vn) - np.array(v)
    np.testing.assert_allclose(a.cpu().reshape(-1).numpy(), np.cumsum(d[::-1])[::-1], atol=1e-5)
    a, _ = pkg.generalized_advantage_estimate(0.99, 0.98, f(v), f(vn), f(r), b([False, True, False]), b([False] * 3))
    d = np.array(r) + 0.99 * np.array(vn) - np.array(v)
    got = a.cpu().reshape(-1).numpy()
    assert abs(got[1] - d[1]) < 1e-5 and abs(got[0] - (d[0] + 0.99 * 0.98 * d[1])) < 1e-5


def test_empty_rollout():
    z = torch.zeros(0, 16, 1, device=DEV)
    a, t = pkg.calculate_advantages(z, z, z, torch.zeros(0, 16, dtype=torch.bool, device=DEV), 0.99, 0.98)
    assert a.shape == (0, 16, 1) and t.shape == (0, 16, 1)


@pytest.mark.parametrize("n,t,norm", [(4096, 128, False), (4096, 128, True), (1024, 128, True), (16, 256, False),
                                      (1, 2048, False), (65536, 1024, False)])
def test_full_size_configs(n, t, norm):
    """BASELINE.json configs at full size: oracle comparison on a row sample + linearity in the reward."""
    roll = O.synthetic_rollout(n, t, 1, 1, seed=1234 + n)
    dev = {k: roll[k].to(DEV) for k in ("reward", "current_state_value", "next_state_value", "terminated")}
    adv, tgt = pkg.calculate_advantages(dev["reward"], dev["current_state_value"], dev["next_state_value"],
                                        dev["terminated"], 0.99, 0.98, normalize_advantage=norm, advantage_scaler=1.0)
    rows = torch.linspace(0, n - 1, min(n, 512)).long().unique()
    a_ref, t_ref = O.calculate_advantages(roll["reward"][rows], roll["current_state_value"][rows],
                                          roll["next_state_value"][rows], roll["terminated"][rows], 0.99, 0.98,
                                          normalize_advantage=norm)
    assert_close(adv[rows.to(DEV)], a_ref, RTOL_FP32, "advantage")
    assert_close(tgt[rows.to(DEV)], t_ref, RTOL_FP32, "target")
    if not norm:  # A is affine in (r, V, V'): A(r1 + r2, V, V') - A(r2, V, V') == A(r1, 0, 0)
        r2 = torch.randn_like(dev["reward"])
        zero = torch.zeros_like(dev["reward"])
        a12, _ = pkg.calculate_advantages(dev["reward"] + r2, dev["current_state_value"], dev["next_state_value"],
                                          dev["terminated"], 0.99, 0.98)
        a2, _ = pkg.calculate_advantages(r2, dev["current_state_value"], dev["next_state_value"], dev["terminated"],
                                         0.99, 0.98)
        a1, _ = pkg.calculate_advantages(dev["reward"], zero, zero, dev["terminated"], 0.99, 0.98)
        assert_close(a12 - a2, a1, 1e-4, "linearity")  # difference of two fp32 results: one extra rounding level
