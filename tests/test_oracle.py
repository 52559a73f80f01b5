"""CPU tests: the oracle restatement against (i) the reference's own outputs (tests/golden, written by
oracle/make_golden.py from /root/reference), (ii) the independent float64 numpy restatement, and
(iii) analytic known-answer cases for the (unpinned, third-party) torchrl GAE recurrence."""
import numpy as np
import pytest
import torch

from oracle import naive
from oracle import ppo_oracle as O
from tests._util import assert_close, load_golden, rel_err, sub

ADV_CASES = ["adv_plain", "adv_norm", "adv_rnorm", "adv_f64"]
TRAIN_CASES = ["train_tanh64", "train_relu3", "train_tanh96"]


@pytest.mark.parametrize("name", ADV_CASES)
def test_calculate_advantages_matches_reference(name):
    g = load_golden(name)
    gamma, lmbda, nr, na, sc = g["cfg"]
    adv, tgt = O.calculate_advantages(torch.from_numpy(g["in_reward"]), torch.from_numpy(g["in_current_state_value"]),
                                      torch.from_numpy(g["in_next_state_value"]), torch.from_numpy(g["in_terminated"]),
                                      gamma, lmbda, bool(nr), bool(na), sc)
    assert adv.dtype == torch.float32 and tuple(adv.shape) == g["advantage"].shape
    assert np.array_equal(adv.numpy(), g["advantage"])
    assert np.array_equal(tgt.numpy(), g["value_target"])


@pytest.mark.parametrize("name", ADV_CASES)
def test_gae_matches_naive_float64(name):
    g = load_golden(name)
    gamma, lmbda, nr, na, sc = g["cfg"]
    r = g["in_reward"][..., 0].astype(np.float64)
    if nr:
        r = naive.normalize_rows_naive(r, sc)
    term = g["in_terminated"]
    done = term.copy()
    done[:, -1] = True
    adv, tgt = naive.gae_naive(r, g["in_current_state_value"][..., 0], g["in_next_state_value"][..., 0], done, term,
                               gamma, lmbda)
    if na:
        adv, tgt = naive.normalize_rows_naive(adv, sc), naive.normalize_rows_naive(tgt, sc)
    assert_close(g["advantage"][..., 0], adv, 1e-5, "advantage")
    assert_close(g["value_target"][..., 0], tgt, 1e-5, "target")


def _gae(r, v, vn, done, term, gamma=0.99, lmbda=0.98):
    f = lambda x: torch.tensor(x, dtype=torch.float32).reshape(1, -1, 1)
    b = lambda x: torch.tensor(x, dtype=torch.bool).reshape(1, -1, 1)
    a, t = O.generalized_advantage_estimate(gamma, lmbda, f(v), f(vn), f(r), b(done), b(term))
    return a.reshape(-1).double().numpy(), t.reshape(-1).double().numpy()


def test_gae_known_answers():
    # T = 1: A = r + gamma*V' - V
    a, t = _gae([2.0], [0.5], [1.0], [True], [False])
    assert abs(a[0] - (2.0 + 0.99 * 1.0 - 0.5)) < 1e-6 and abs(t[0] - (a[0] + 0.5)) < 1e-6
    # every step terminated: A = r - V (no bootstrap, no carry)
    r, v, vn = [1.0, -2.0, 3.0], [0.3, 0.2, 0.1], [9.0, 9.0, 9.0]
    a, _ = _gae(r, v, vn, [True] * 3, [True] * 3)
    np.testing.assert_allclose(a, np.array(r) - np.array(v), atol=1e-6)
    # lmbda = 0: A = delta
    a, _ = _gae(r, v, vn, [False] * 3, [False] * 3, lmbda=0.0)
    np.testing.assert_allclose(a, np.array(r) + 0.99 * np.array(vn) - np.array(v), atol=1e-5)
    # gamma = lmbda = 1, nothing ends: suffix sums of delta
    a, _ = _gae(r, v, vn, [False] * 3, [False] * 3, gamma=1.0, lmbda=1.0)
    d = np.array(r) + np.array(vn) - np.array(v)
    np.testing.assert_allclose(a, np.cumsum(d[::-1])[::-1], atol=1e-5)
    # done without terminated (the forced last-step done, ppo.py:72): bootstrap kept, carry cut
    a, _ = _gae(r, v, vn, [False, True, False], [False, False, False])
    d = np.array(r) + 0.99 * np.array(vn) - np.array(v)
    assert abs(a[1] - d[1]) < 1e-5 and abs(a[0] - (d[0] + 0.99 * 0.98 * d[1])) < 1e-5


def test_gae_shape_mismatch_raises():
    x = torch.zeros(2, 5, 1)
    with pytest.raises(RuntimeError):
        O.generalized_advantage_estimate(0.99, 0.98, x, x, x[:, :4], x.bool(), x.bool())


def test_gae_time_dim():
    g = torch.Generator().manual_seed(0)
    v, vn, r = (torch.randn(3, 7, 1, generator=g) for _ in range(3))
    d = torch.rand(3, 7, 1, generator=g) < 0.2
    a0, t0 = O.generalized_advantage_estimate(0.9, 0.8, v, vn, r, d, d)
    tr = lambda x: x.transpose(0, 1).contiguous()
    a1, t1 = O.generalized_advantage_estimate(0.9, 0.8, tr(v), tr(vn), tr(r), tr(d), tr(d), time_dim=0)
    assert torch.equal(tr(a0), a1) and torch.equal(tr(t0), t1)


def _agent_from_golden(g):
    B, epochs, lr, clip, ent, out_max = g["cfg"]
    init = sub(g, "init/")
    D = init["actor.actor.first_layers.0.weight"].shape[1]
    A = init["actor.actor_logstd"].shape[0]
    cfg = O.OracleConfig(obs_dim=D, act_dim=A, actor_hidden=[int(h) for h in g["hidden"]], critic_hidden=[128, 128],
                         activation=str(g["activation"]), output_max_value=float(out_max), learning_rate=float(lr),
                         batch_size=int(B), epochs=int(epochs), clip_epsilon=float(clip), entropy_eps=float(ent))
    agent = O.OracleAgent(cfg)
    agent.networks.load_state_dict({k: torch.from_numpy(v) for k, v in init.items()})
    return agent, cfg


def _flat_mem(g):
    mem = sub(g, "mem/")
    M = mem["action_log_prob"].size
    return {
        "current_state": torch.from_numpy(mem["current_state"]).reshape(M, -1),
        "action": torch.from_numpy(mem["action"]).reshape(M, -1),
        "action_log_prob": torch.from_numpy(mem["action_log_prob"]).reshape(M),
        "advantage": torch.from_numpy(mem["advantage"]).reshape(M, 1),
        "current_state_value_target": torch.from_numpy(mem["current_state_value_target"]).reshape(M, 1),
    }


@pytest.mark.parametrize("name", TRAIN_CASES)
def test_train_matches_reference(name):
    torch.set_num_threads(1)
    g = load_golden(name)
    agent, cfg = _agent_from_golden(g)
    losses = O.ppo_train(agent, _flat_mem(g), [torch.from_numpy(p) for p in g["perms"]])
    final = sub(g, "final/")
    for k, v in agent.networks.state_dict().items():
        assert_close(v, final[k], 1e-6, k)
    for oname, opt in agent.optimizers.items():
        for pid, st in opt.state_dict()["state"].items():
            assert_close(st["exp_avg"], g[f"opt/{oname}/{pid}/exp_avg"], 1e-6, f"{oname}/{pid}/exp_avg")
            assert_close(st["exp_avg_sq"], g[f"opt/{oname}/{pid}/exp_avg_sq"], 1e-6, f"{oname}/{pid}/exp_avg_sq")
            assert float(st["step"]) == float(g[f"opt/{oname}/{pid}/step"])
    nb = len(losses) // int(cfg.epochs)
    ep = np.array(losses).reshape(int(cfg.epochs), nb, 2).mean(axis=1).mean(axis=0)
    np.testing.assert_allclose(ep, g["logged_losses"], rtol=1e-5)


@pytest.mark.parametrize("name", TRAIN_CASES)
def test_minibatch_grads_match_naive_float64(name):
    g = load_golden(name)
    agent, cfg = _agent_from_golden(g)
    mem = _flat_mem(g)
    idx = torch.from_numpy(g["perms"][0])[:cfg.batch_size]
    obs, act = mem["current_state"][idx], mem["action"][idx]
    oldlp, adv, tgt = mem["action_log_prob"][idx], mem["advantage"][idx], mem["current_state_value_target"][idx]
    al, cl, grads, logp, value = O.minibatch_grads(agent, obs, act, oldlp, adv, tgt)
    sd = {k: v.detach().numpy() for k, v in agent.networks.state_dict().items()}

    def layers(prefix, n_hidden):
        out = [(sd[f"{prefix}.first_layers.{2 * i}.weight"], sd[f"{prefix}.first_layers.{2 * i}.bias"])
               for i in range(n_hidden)]
        return out + [(sd[f"{prefix}.last_layer.weight"], sd[f"{prefix}.last_layer.bias"])]

    aw, cw = layers("actor.actor", len(cfg.actor_hidden)), layers("critic.network", 2)
    res = naive.ppo_minibatch_naive(aw, sd["actor.actor_logstd"], cw, obs.numpy(), act.numpy(), oldlp.numpy(),
                                    adv.numpy(), tgt.numpy(), cfg.activation, cfg.clip_epsilon, cfg.entropy_eps,
                                    cfg.output_max_value)
    assert abs(al - res["actor_loss"]) <= 1e-5 * max(1.0, abs(res["actor_loss"]))
    assert abs(cl - res["critic_loss"]) <= 1e-5 * max(1.0, abs(res["critic_loss"]))
    assert_close(logp, res["logp"], 1e-5, "logp")
    assert_close(value, res["value"], 1e-5, "value")
    for i, (dW, db) in enumerate(res["actor_grads"]):
        key = f"actor.actor.first_layers.{2 * i}" if i < len(cfg.actor_hidden) else "actor.actor.last_layer"
        assert_close(grads[key + ".weight"], dW, 1e-5, key)
        assert_close(grads[key + ".bias"], db, 1e-5, key)
    assert_close(grads["actor.actor_logstd"], res["logstd_grad"], 1e-5, "logstd")
    for i, (dW, db) in enumerate(res["critic_grads"]):
        key = f"critic.network.first_layers.{2 * i}" if i < 2 else "critic.network.last_layer"
        assert_close(grads[key + ".weight"], dW, 1e-5, key)
        assert_close(grads[key + ".bias"], db, 1e-5, key)


def test_adam_naive_matches_torch():
    g = torch.Generator().manual_seed(3)
    p = torch.randn(257, generator=g)
    p0 = p.clone()
    p.requires_grad_(True)
    opt = torch.optim.Adam([p], lr=3e-4, foreach=False)
    pn, m, v = p0.double().numpy(), np.zeros(257), np.zeros(257)
    for step in range(1, 6):
        gr = torch.randn(257, generator=g)
        p.grad = gr.clone()
        opt.step()
        pn, m, v = naive.adam_step_naive(pn, gr.double().numpy(), m, v, step, 3e-4)
    assert rel_err(p.detach() - p0, pn - p0.double().numpy()) < 1e-3  # the update itself (fp32 cancellation on p - p0)
    assert_close(p, pn, 1e-6, "params")


def test_normalize_state_matches_reference():
    g = load_golden("obsnorm")
    out = O.normalize_state(torch.from_numpy(g["in_observation"]))
    assert out.dtype == torch.float32 and tuple(out.shape) == g["state"].shape
    assert np.array_equal(out.numpy(), g["state"])
