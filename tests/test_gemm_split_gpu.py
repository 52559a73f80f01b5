"""fp32-tolerance GEMMs on the tensor cores (csrc/gemm_split.cu) against float64.

Every fp32 operand value goes in as three bf16 terms and six bf16 MMAs per product accumulate in fp32 TMEM.  The tensor
core truncates that accumulator after every K = 16 step (profiles/rz_probe.py), which is why the kernel runs the small
products first and the callers keep one accumulator to 512 samples of a weight gradient; what is left is about half an
ulp per step of the leading product.  Tolerance: 3e-6 x max |C| per element (north_star's fp32 bound is 1e-5 on the whole
update; torch's fp32 matmul measures 2e-7 ... 9e-7 on the same inputs).
"""
import pytest
import torch

from mujoco_reinforcement_learning_b200 import _lib

pytestmark = pytest.mark.gpu
DEV = "cuda"


def run(A, B, M, N, K, a_mn, b_mn, split=1, bias=False):
    lib = _lib.load()
    C = torch.empty(M, N, device=DEV)
    bg = torch.empty(M, device=DEV) if bias else None
    _lib.check(lib.b200ppo_debug_gemm_split(_lib.ptr(A), _lib.ptr(B), _lib.ptr(C), _lib.ptr(bg) if bias else None, M, N, K,
                                            a_mn, b_mn, split, _lib.stream_ptr()), "debug_gemm_split")
    return C, bg


CASES = [
    # M, N, K, a_mn, b_mn, split
    (4096, 256, 376, 0, 0, 1),    # forward, first layer (K tail: 376 = 5 x 64 + 56)
    (4096, 256, 256, 0, 0, 1),    # forward, second layer
    (1000, 256, 256, 0, 1, 1),    # dgrad (weights read MN-major), ragged M
    (300, 376, 256, 0, 1, 1),     # dgrad to the observations
    (256, 376, 4096, 1, 1, 8),    # weight gradient, split-K
    (256, 256, 5000, 1, 1, 11),   # K not a multiple of the tile, uneven splits
    (17, 256, 4096, 1, 1, 8),     # output layer's weight gradient (M < one tile)
    (1, 256, 2048, 1, 1, 40),     # critic head; more splits than k tiles per product
    (130, 70, 33, 0, 0, 1),
]


@pytest.mark.parametrize("terms", [3, 2], ids=["bf16x3", "fp16x2"])
@pytest.mark.parametrize("M,N,K,a_mn,b_mn,split", CASES)
def test_gemm_split_matches_float64(M, N, K, a_mn, b_mn, split, terms, monkeypatch):
    monkeypatch.setenv("B200PPO_DEBUG_SPLIT_TERMS", str(terms))
    g = torch.Generator().manual_seed(M * 7 + N * 3 + K)
    A = torch.randn(M, K, generator=g) * torch.exp(torch.randn(M, 1, generator=g))  # rows of very different scale
    B = torch.randn(N, K, generator=g)
    ref = (A.double() @ B.double().T)
    Ad = (A.T.contiguous() if a_mn else A).to(DEV)
    Bd = (B.T.contiguous() if b_mn else B).to(DEV)
    bias = bool(a_mn and b_mn)
    C, bg = run(Ad, Bd, M, N, K, a_mn, b_mn, split, bias)
    err = (C.cpu().double() - ref).abs().max().item() / ref.abs().max().item()
    print(f"max err / max |C| = {err:.2e}")
    assert err < 3e-6, err
    if bias:
        ref_b = A.double().sum(1)
        # a sum of K values: compare against the sum of magnitudes (what fp32 accumulation is bounded by)
        errb = ((bg.cpu().double() - ref_b).abs() / A.double().abs().sum(1)).max().item()
        assert errb < 3e-6, errb


@pytest.mark.parametrize("terms", [3, 2], ids=["bf16x3", "fp16x2"])
def test_gemm_split_keeps_small_values(terms, monkeypatch):
    """Values 2^-20 below their row's scale survive (a plain bf16 operand would drop them)."""
    monkeypatch.setenv("B200PPO_DEBUG_SPLIT_TERMS", str(terms))
    M, N, K = 128, 128, 64
    A = torch.ones(M, K)
    A[:, 1] = 2.0 ** -20
    B = torch.zeros(N, K)
    B[:, 1] = 1.0
    C, _ = run(A.to(DEV), B.to(DEV), M, N, K, 0, 0)
    assert torch.equal(C.cpu(), torch.full((M, N), 2.0 ** -20))


@pytest.mark.parametrize("shape", [dict(D=376, A=17, B=4096), dict(D=376, A=17, B=2500, act="relu"), dict(D=27, A=8, B=8192),
                                   dict(D=376, A=17, B=32768)],  # the bench minibatch: 64 split-K partials of 512 samples each
                         ids=lambda s: "-".join(f"{k}{v}" for k, v in s.items()))
@pytest.mark.parametrize("terms", [3, 2], ids=["bf16x3", "fp16x2"])
def test_fp32_minibatch_on_tensor_cores_vs_oracle(shape, terms):
    """The fp32-precision minibatch at sizes where its GEMMs take the tensor-core route: losses and every gradient at
    north_star's fp32 tolerance (1e-5 of the tensor's scale) against the oracle's autograd (ppo.py:110-134)."""
    from tests._util import RTOL_FP32, assert_close
    from tests.test_chain_gpu import reference_graph
    from tests.test_update_gpu import make_pair
    D, A, B = shape["D"], shape["A"], shape["B"]
    oracle, agent, run = make_pair(D, A, [256, 256], [256, 256], shape.get("act", "tanh"), batch=B, max_batch=B, precision="fp32", seed=3)
    g = torch.Generator().manual_seed(9)
    obs = torch.randn(B, D, generator=g)
    action = torch.randn(B, A, generator=g).clamp_(-3, 3)
    adv = torch.randn(B, 1, generator=g)
    tgt = torch.randn(B, 1, generator=g)
    with torch.no_grad():
        for n, p in oracle.networks.named_parameters():
            if n.endswith("bias"):
                p.add_(0.05 * torch.randn(p.shape, generator=g))
        agent.networks.load_state_dict(oracle.networks.state_dict())
        mean, std = oracle.networks["actor"](obs)
        old_logp = torch.distributions.Normal(mean, std).log_prob(action).sum(1) + 0.08 * torch.randn(B, generator=g)
    _, grads_ref, al, cl = reference_graph(oracle, obs, action, old_logp, adv, tgt)
    eng = agent.engine
    eng.set_fp32_terms(terms)
    hp = eng.hparams(1e-4, 1e-4, oracle.cfg.clip_epsilon, oracle.cfg.entropy_eps)
    losses, grads = eng.minibatch_grads(obs.to(DEV), action.to(DEV), old_logp.to(DEV), adv.to(DEV), tgt.to(DEV), hp)
    assert abs(losses[0].item() - al) <= RTOL_FP32 * max(1.0, abs(al)), (losses[0].item(), al)
    assert abs(losses[1].item() - cl) <= RTOL_FP32 * max(1.0, abs(cl)), (losses[1].item(), cl)
    by_name = eng.grads_by_name(grads, agent.networks.named_parameters())
    for k, r in grads_ref.items():
        assert_close(by_name[k], r, RTOL_FP32, f"grad {k}")


@pytest.mark.parametrize("hidden,critic_hidden,D,B", [([256, 192, 128], [128, 64], 200, 4096), ([320, 256], [256, 256], 200, 4096),
                                                      ([250, 130], [250, 130], 377, 4097)],
                         ids=["3-layer-actor", "wide-first-layer", "ragged-everything"])
@pytest.mark.parametrize("terms", [3, 2], ids=["bf16x3", "fp16x2"])
def test_fp32_tensor_core_route_on_other_architectures(hidden, critic_hidden, D, B, terms):
    """Shapes that mix the routes inside one minibatch: 128-wide layers (the one-tile kernel in three-term mode), a
    weight-gradient group with more problems than one tensor-core launch takes (falls back to the FFMA kernel), layers
    below the size threshold — losses and every gradient still at 1e-5 against the oracle's autograd."""
    from tests._util import RTOL_FP32, assert_close
    from tests.test_update_gpu import make_pair
    A = 6  # ragged-everything: widths that are no multiple of 16, an odd observation width (scalar split loads), one row past a tile
    oracle, agent, run = make_pair(D, A, hidden, critic_hidden, "tanh", batch=B, max_batch=B, precision="fp32", seed=5)
    g = torch.Generator().manual_seed(13)
    obs = torch.randn(B, D, generator=g)
    action = torch.randn(B, A, generator=g).clamp_(-3, 3)
    adv = torch.randn(B, 1, generator=g)
    tgt = torch.randn(B, 1, generator=g)
    with torch.no_grad():
        mean, std = oracle.networks["actor"](obs)
        old_logp = torch.distributions.Normal(mean, std).log_prob(action).sum(1) + 0.05 * torch.randn(B, generator=g)
        v_ref = oracle.networks["critic"](obs)
    # losses and gradients through autograd on the oracle's own modules (ppo.py:110-134)
    cfg = oracle.cfg
    for p in oracle.networks.parameters():
        p.grad = None
    m2, s2 = oracle.networks["actor"](obs)
    dist = torch.distributions.Normal(m2, s2)
    ratio = (dist.log_prob(action).sum(1) - old_logp).exp()[:, None]
    s1, s2_ = ratio * adv, torch.clamp(ratio, 1.0 - cfg.clip_epsilon, 1.0 + cfg.clip_epsilon) * adv
    actor_loss = -torch.min(s1, s2_).mean() - dist.entropy().mean() * cfg.entropy_eps
    from oracle import ppo_oracle as O
    critic_loss = O.huber_loss(oracle.networks["critic"](obs), tgt, reduction="mean")
    (actor_loss + critic_loss).backward()
    grads_ref = {n: p.grad.detach().clone() for n, p in oracle.networks.named_parameters()}
    eng = agent.engine
    eng.set_fp32_terms(terms)
    hp = eng.hparams(1e-4, 1e-4, cfg.clip_epsilon, cfg.entropy_eps)
    losses, grads = eng.minibatch_grads(obs.to(DEV), action.to(DEV), old_logp.to(DEV), adv.to(DEV), tgt.to(DEV), hp)
    assert abs(losses[0].item() - actor_loss.item()) <= RTOL_FP32 * max(1.0, abs(actor_loss.item()))
    assert abs(losses[1].item() - critic_loss.item()) <= RTOL_FP32 * max(1.0, abs(critic_loss.item()))
    by_name = eng.grads_by_name(grads, agent.networks.named_parameters())
    for k, r in grads_ref.items():
        assert_close(by_name[k], r, RTOL_FP32, f"grad {k}")
    # and the forward-only entry point on the same context
    action_o, logp_o, value_o, mean_o = eng.policy_infer(obs.to(DEV), None)
    assert_close(mean_o, mean, RTOL_FP32, "mean")
    assert_close(value_o, v_ref, RTOL_FP32, "value")
