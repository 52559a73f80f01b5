"""SAC update step on the GPU (sac.py + library kernels) against the reference's own outputs (tests/golden/ref_sac_*)
and against the CPU oracle on a larger seeded case.  fp32 path: 1e-5 relative to each tensor's scale."""
import numpy as np
import pytest
import torch

import mujoco_reinforcement_learning_b200 as pkg
from oracle import sac_oracle as S
from tests._util import assert_close, assert_params_close, load_golden, sub
from tests.test_sac_oracle import SAC_CASES, flat_replay, sac_agent_from_golden

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _run_from_cfg(cfg: S.SacConfig, obs_dim: int, window: int):
    act = {"tanh": torch.nn.Tanh, "relu": torch.nn.ReLU}[cfg.activation]
    return pkg.Run(training_config=pkg.TrainingConfig(learning_rate=cfg.learning_rate, batch_size=cfg.batch_size),
                   ppo_config=pkg.PPOConfig(max_grad_norm=cfg.max_grad_norm),
                   sac_config=pkg.SACConfig(gamma=cfg.gamma, alpha=cfg.alpha, tau=cfg.tau, target_update_interval=cfg.target_update_interval),
                   environment_config=pkg.EnvironmentConfig(window_length=window),
                   network_config=pkg.NetworkConfig(input_shape=obs_dim, output_shape=cfg.act_dim, output_max_value=cfg.output_max_value,
                                                    activation_class=act, num_linear_layers=len(cfg.hidden),
                                                    linear_hidden_shapes=list(cfg.hidden)),
                   device=DEV)


class _Helper:
    def __init__(self, run):
        self.run = run


@pytest.mark.parametrize("name", SAC_CASES)
def test_sac_train_matches_reference_golden(name):
    g = load_golden(name)
    oracle_agent, cfg = sac_agent_from_golden(g)
    mem = sub(g, "mem/")
    window, obs_dim = mem["current_state"].shape[2], mem["current_state"].shape[3]
    run = _run_from_cfg(cfg, obs_dim, window)
    agent = pkg.SoftActorCriticAgent(run)
    algo = pkg.SoftActorCritic(_Helper(run), agent)
    agent.networks.load_state_dict({k: torch.from_numpy(v) for k, v in sub(g, "init/").items()})
    memory = {k: torch.from_numpy(v).to(DEV) for k, v in mem.items()}
    perms, eps = torch.from_numpy(g["perms"]), torch.from_numpy(g["eps"]).to(DEV)
    for call in range(perms.shape[0]):
        losses = algo.train(memory, call, idx=perms[call], noise=eps[2 * call:2 * call + 2])
        assert_close(losses[:4].cpu(), g["losses"][call][:4], 1e-5, f"{name} losses of call {call}")
    final = sub(g, "final/")
    for k, v in agent.networks.state_dict().items():
        assert_close(v.cpu(), final[k], 1e-5, f"{name} {k}")


def test_sac_step_matches_oracle_larger_case():
    torch.manual_seed(5)
    cfg = S.SacConfig(state_dim=2 * 24, act_dim=6, hidden=[128, 96], activation="tanh", learning_rate=3e-4, batch_size=1024,
                      alpha=0.2, tau=0.01)
    oracle_agent = S.OracleSacAgent(cfg)
    run = _run_from_cfg(cfg, 24, 2)
    agent = pkg.SoftActorCriticAgent(run)
    algo = pkg.SoftActorCritic(_Helper(run), agent)
    agent.networks.load_state_dict(oracle_agent.networks.state_dict())
    replay = S.synthetic_replay(32, 64, (2, 24), 6, seed=77)
    flat = {k: v.reshape(32 * 64, *v.shape[2:]) for k, v in replay.items()}
    memory = {k: v.to(DEV) for k, v in replay.items()}
    g = torch.Generator().manual_seed(9)
    for call in range(3):
        idx = torch.randperm(32 * 64, generator=g)
        eps = torch.randn(2, cfg.batch_size, cfg.act_dim, generator=g)
        ref = S.sac_train_step(oracle_agent, flat, idx, eps[0], eps[1], call)
        got = algo.train(memory, call, idx=idx, noise=eps.to(DEV))
        assert_close(got[:4].cpu(), torch.tensor(ref[:4]), 1e-5, f"losses of call {call}")
    ref_sd = oracle_agent.networks.state_dict()
    for k, v in agent.networks.state_dict().items():
        assert_close(v.cpu(), ref_sd[k], 2e-5, k)


def test_polyak_update_is_bit_exact():
    g = torch.Generator(device=DEV).manual_seed(3)
    t, s = torch.randn(100003, device=DEV, generator=g), torch.randn(100003, device=DEV, generator=g)
    ref = t * (1.0 - 0.005) + s * 0.005
    pkg.polyak_update_(t, s, 0.005)
    assert torch.equal(t, ref)


def test_sac_checkpoint_round_trip_and_lr_schedule(tmp_path):
    """the SAC agent keeps the reference's schedulers (ExponentialLR(0.999) per optimiser,
    soft_actor_critic_agent.py:30-34) and `Agent.save` / `load` files (agent.py:47-72) in torch's Adam state layout."""
    g = load_golden(SAC_CASES[0])
    _, cfg = sac_agent_from_golden(g)
    mem = sub(g, "mem/")
    window, obs_dim = mem["current_state"].shape[2], mem["current_state"].shape[3]
    run = _run_from_cfg(cfg, obs_dim, window)
    run.experiment_path = str(tmp_path)
    agent = pkg.SoftActorCriticAgent(run)
    algo = pkg.SoftActorCritic(_Helper(run), agent)
    memory = {k: torch.from_numpy(v).to(DEV) for k, v in mem.items()}
    perms, eps = torch.from_numpy(g["perms"]), torch.from_numpy(g["eps"]).to(DEV)
    algo.train(memory, 0, idx=perms[0], noise=eps[0:2])
    lr0 = agent.optimizers["actor"].lr
    for sch in agent.schedulers.values():
        sch.step()
    assert agent.optimizers["actor"].lr == pytest.approx(lr0 * 0.999) and set(agent.schedulers) == {"actor", "online_critic"}
    agent.save()
    d = tmp_path / "networks" / "0"
    assert sorted(p.name for p in d.iterdir()) == ["networks.pth", "optimizer_actor.pth", "optimizer_online_critic.pth"]
    # the optimizer file loads into a stock torch.optim.Adam over same-shaped parameters
    osd = torch.load(d / "optimizer_actor.pth", map_location="cpu")
    shapes = [st["exp_avg"].shape for st in osd["state"].values()]
    stock = torch.optim.Adam([torch.nn.Parameter(torch.zeros(s)) for s in shapes], lr=1.0)
    stock.load_state_dict(osd)
    assert stock.param_groups[0]["lr"] == pytest.approx(lr0 * 0.999)
    # and back into a fresh agent
    run2 = _run_from_cfg(cfg, obs_dim, window)
    run2.experiment_path = str(tmp_path)
    agent2 = pkg.SoftActorCriticAgent(run2)
    agent2.load()
    for k, v in agent.networks.state_dict().items():
        assert torch.equal(v, agent2.networks.state_dict()[k]), k
    assert agent2.optimizers["online_critic"].step_count == 1
    assert torch.equal(agent2.engine_q.exp_avg, agent.engine_q.exp_avg)
