"""K2 (bit-exact gather) and K5 (Adam) parity through the C ABI."""
import numpy as np
import pytest
import torch

import mujoco_reinforcement_learning_b200 as pkg
from tests._util import RTOL_FP32, assert_close

pytestmark = pytest.mark.gpu
DEV = "cuda"


@pytest.mark.parametrize("m,d,a", [(1, 1, 1), (100, 376, 17), (257, 27, 8), (64, 11, 3), (1000, 17, 6), (33, 128, 32)])
def test_gather_minibatch_bit_exact(m, d, a):
    g = torch.Generator().manual_seed(m + d)
    obs, act = torch.randn(m, d, generator=g), torch.randn(m, a, generator=g)
    lp, adv, tgt = (torch.randn(m, generator=g) for _ in range(3))
    idx = torch.randperm(m, generator=g)
    out = pkg.gather_minibatch(idx.to(DEV), obs.to(DEV), act.to(DEV), lp.to(DEV), adv.to(DEV), tgt.to(DEV))
    for got, ref in zip(out, (obs, act, lp, adv, tgt)):
        assert torch.equal(got.cpu(), ref[idx])  # bit-exact
    part = idx[: m // 2]
    out = pkg.gather_minibatch(part.to(DEV), obs.to(DEV), act.to(DEV), lp.to(DEV), adv.to(DEV), tgt.to(DEV))
    assert torch.equal(out[0].cpu(), obs[part]) and out[0].shape[0] == m // 2


def test_gather_full_humanoid_epoch_is_a_permutation():
    n, d, a = 4096 * 128, 376, 17
    g = torch.Generator(device=DEV).manual_seed(5)
    obs = torch.randn(n, d, device=DEV, generator=g)
    act = torch.randn(n, a, device=DEV, generator=g)
    s = torch.randn(n, device=DEV, generator=g)
    idx = torch.randperm(n, device=DEV, generator=g)
    o, ac, lp, _, _ = pkg.gather_minibatch(idx, obs, act, s, s, s)
    assert torch.equal(o, obs[idx]) and torch.equal(ac, act[idx]) and torch.equal(lp, s[idx])
    # checksum of checksums: a permutation preserves the multiset of row sums
    assert torch.equal(torch.sort(o.double().sum(1)).values, torch.sort(obs.double().sum(1)).values)


def test_gather_rows_any_dtype_negative_and_out_of_range():
    g = torch.Generator().manual_seed(1)
    for src in (torch.randn(50, 3, 5, generator=g), torch.randn(50, generator=g).double(),
                torch.rand(50, 7, generator=g) < 0.5, torch.randint(0, 255, (50, 3), generator=g, dtype=torch.uint8)):
        idx = torch.tensor([0, 49, -1, -50, 7, 7])
        assert torch.equal(pkg.gather_rows(src.to(DEV), idx.to(DEV)).cpu(), src[idx])
    with pytest.raises(IndexError):
        pkg.gather_rows(torch.zeros(4, 2, device=DEV), torch.tensor([4], device=DEV))
    e = pkg.gather_rows(torch.zeros(4, 2, device=DEV), torch.zeros(0, dtype=torch.int64, device=DEV))
    assert e.shape == (0, 2)


def test_rollout_memory_index_gathers_every_leaf():
    g = torch.Generator().manual_seed(2)
    leaves = {"current_state": torch.randn(6, 5, 1, 12, generator=g), "reward": torch.randn(6, 5, 1, generator=g).double(),
              "terminated": torch.rand(6, 5, generator=g) < 0.3, "action_log_prob": torch.randn(6, 5, generator=g)}
    mem = pkg.RolloutMemory({k: v.to(DEV) for k, v in leaves.items()}, (6, 5)).view(-1)
    idx = torch.randperm(30, generator=g)
    sh = mem[idx.to(DEV)]
    for k, v in leaves.items():
        assert torch.equal(sh[k].cpu(), v.reshape(30, *v.shape[2:])[idx]), k
    assert len(sh[4:14]) == 10


@pytest.mark.parametrize("n", [1, 3, 1024, 329251])
def test_adam_matches_torch_single_tensor(n):
    g = torch.Generator().manual_seed(n)
    p0 = torch.randn(n, generator=g)
    ref = p0.clone().requires_grad_(True)
    opt = torch.optim.Adam([ref], lr=1e-4, foreach=False)
    p, m, v = p0.clone().to(DEV), torch.zeros(n, device=DEV), torch.zeros(n, device=DEV)
    for step in range(1, 7):
        gr = torch.randn(n, generator=g) * (10.0 ** float(torch.randint(-4, 2, (1,), generator=g)))
        ref.grad = gr.clone()
        opt.step()
        pkg.adam_step_(p, gr.to(DEV), m, v, step, 1e-4)
    st = opt.state[ref]
    assert_close(p, ref.detach(), 1e-6, "param")
    assert_close(p.cpu() - p0, ref.detach() - p0, 2e-3, "update")  # p - p0 cancels ~3 digits in fp32
    assert_close(m, st["exp_avg"], RTOL_FP32, "exp_avg")
    assert_close(v, st["exp_avg_sq"], RTOL_FP32, "exp_avg_sq")
