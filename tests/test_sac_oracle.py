"""The SAC oracle (oracle/sac_oracle.py) against outputs of the reference's own `SoftActorCritic.train`
(tests/golden/ref_sac_*.npz, written by oracle/make_golden.py running the reference verbatim)."""
import numpy as np
import pytest
import torch

from oracle import sac_oracle as S
from tests._util import assert_close, load_golden, sub

SAC_CASES = ["sac_tanh32", "sac_relu48"]


def sac_agent_from_golden(g):
    B, lr, gamma, alpha, tau, interval, max_norm, out_max, W = g["cfg"].tolist()
    mem = sub(g, "mem/")
    state_dim = int(np.prod(mem["current_state"].shape[2:]))
    cfg = S.SacConfig(state_dim=state_dim, act_dim=mem["action"].shape[-1], hidden=[int(h) for h in g["hidden"]],
                      activation=str(g["activation"]), output_max_value=out_max, learning_rate=lr, batch_size=int(B),
                      gamma=gamma, alpha=alpha, tau=tau, target_update_interval=int(interval), max_grad_norm=max_norm)
    agent = S.OracleSacAgent(cfg)
    agent.networks.load_state_dict({k: torch.from_numpy(v) for k, v in sub(g, "init/").items()})
    return agent, cfg


def flat_replay(g):
    mem = sub(g, "mem/")
    n, t = mem["reward"].shape[:2]
    return {k: torch.from_numpy(v).reshape(n * t, *v.shape[2:]) for k, v in mem.items()}


@pytest.mark.parametrize("name", SAC_CASES)
def test_sac_train_matches_reference(name):
    g = load_golden(name)
    agent, cfg = sac_agent_from_golden(g)
    mem = flat_replay(g)
    perms, eps = torch.from_numpy(g["perms"]), torch.from_numpy(g["eps"])
    for call in range(perms.shape[0]):
        losses = S.sac_train_step(agent, mem, perms[call], eps[2 * call], eps[2 * call + 1], call)
        assert_close(torch.tensor(losses[:4]), g["losses"][call][:4], 1e-5, f"{name} losses of call {call}")
    final = sub(g, "final/")
    for k, v in agent.networks.state_dict().items():
        assert_close(v, final[k], 1e-5, f"{name} {k}")
    for oname in ("actor", "online_critic"):
        sd = agent.optimizers[oname].state_dict()["state"]
        for pid, st in sd.items():
            assert_close(st["exp_avg"], g[f"opt/{oname}/{pid}/exp_avg"], 1e-5, f"{name} {oname} exp_avg {pid}")
            assert_close(st["exp_avg_sq"], g[f"opt/{oname}/{pid}/exp_avg_sq"], 1e-5, f"{name} {oname} exp_avg_sq {pid}")


def test_soft_update_known_answer():
    a, b = torch.nn.Linear(3, 2), torch.nn.Linear(3, 2)
    wa, wb = a.weight.detach().clone(), b.weight.detach().clone()
    S.soft_update(a, b, 0.25)
    assert torch.allclose(a.weight, wa * 0.75 + wb * 0.25)
    S.hard_update(a, b)
    assert torch.equal(a.weight, b.weight)
