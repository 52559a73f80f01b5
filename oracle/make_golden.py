"""Generate `tests/golden/ref_*.npz` by running the REFERENCE'S OWN CODE — TEST INFRASTRUCTURE ONLY.

Build-container only:  `python -m oracle.make_golden`  (needs `/root/reference`, read-only).
Each case runs in its own process because the reference's `Run` config is a process-wide
singleton (`src/utils/type_utils.py:1-7`).  What executes, unmodified, from
`/root/reference/src`:

 * `entities.algorithms.ppo.PPO.calculate_advantages` (ppo.py:62-91) and `.train` (ppo.py:93-154)
 * `entities.agents.ppo_agent.PPOAgent` (ppo_agent.py:10-43) — with its two module-level names
   `Actor`/`Critic` re-pointed at the reference's MLP classes `models.linear.actor.Actor` and
   `models.critic.Critic` (as committed it binds the LSTM variants)
 * `models.network_block_creator.NetworkBlock`, `entities.features.Run` and friends.

Third-party packages the reference needs but the image lacks are shimmed (`oracle/shims.py`).
`torch.randperm` is wrapped (not replaced) to record the permutations `train` draws, so the
CUDA path can be fed the same indices.
"""
from __future__ import annotations

import argparse
import os
import subprocess
import sys
import tempfile

import numpy as np
import torch

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")

CASES = {
    # name: obs, act, actor hidden, activation, N, T, batch, epochs, flags
    "adv_plain": dict(kind="adv", N=6, T=37, seed=11),
    "adv_norm": dict(kind="adv", N=5, T=64, seed=12, normalize_advantage=True, advantage_scaler=0.7),
    "adv_rnorm": dict(kind="adv", N=3, T=130, seed=13, normalize_rewards=True, advantage_scaler=1.3),
    "adv_f64": dict(kind="adv", N=4, T=50, seed=14, reward_f64=True),
    "obsnorm": dict(kind="obsnorm", N=5, T=1, seed=31),
    "train_tanh64": dict(kind="train", D=17, A=6, hidden=[64, 64], act="Tanh", N=4, T=64, B=48, epochs=2, seed=21),
    "train_relu3": dict(kind="train", D=11, A=3, hidden=[32, 24, 16], act="ReLU", N=3, T=50, B=32, epochs=2,
                        seed=22),
    "train_tanh96": dict(kind="train", D=27, A=8, hidden=[96, 80], act="Tanh", N=8, T=32, B=64, epochs=1, seed=23,
                         out_max=2.0),
    # Soft Actor-Critic update step: two consecutive SoftActorCritic.train calls on a replay memory
    "sac_tanh32": dict(kind="sac", D=7, W=2, A=3, hidden=[32, 24], act="Tanh", N=6, T=40, B=64, seed=41, calls=2),
    "sac_relu48": dict(kind="sac", D=5, W=2, A=4, hidden=[48, 48], act="ReLU", N=4, T=50, B=96, seed=42, calls=2,
                       out_max=2.0, lr=3e-4),
}


def _build_run(spec, tmpdir):
    from entities import features as F
    act_cls = getattr(torch.nn, spec.get("act", "Tanh"))
    hidden = spec.get("hidden", [8, 8])
    return F.Run(
        F.RewardConfig(),
        F.TrainingConfig(iteration_count=1, learning_rate=spec.get("lr", 1e-4), weight_decay=1e-4,
                         batch_size=spec.get("B", 4), epochs_per_iteration=spec.get("epochs", 1),
                         minimum_learning_rate=1e-4),
        F.PPOConfig(max_grad_norm=1.0, clip_epsilon=spec.get("clip", 0.1), gamma=0.99, lmbda=0.98,
                    entropy_eps=spec.get("ent", 1e-4), advantage_scaler=spec.get("advantage_scaler", 1.0),
                    normalize_advantage=spec.get("normalize_advantage", False), critic_coeffiecient=1.0),
        F.SACConfig(1.0, 0.99, 0.05, 0.005, 999, 1, False),
        F.EnvironmentConfig(maximum_timesteps=spec["T"], num_envs=spec["N"], window_length=spec.get("W", 1)),
        F.AgentConfig(sub_action_count=1),
        F.NetworkConfig(input_shape=spec.get("D", 4), output_shape=spec.get("A", 2),
                        output_max_value=spec.get("out_max", 1.0), activation_class=act_cls,
                        num_linear_layers=len(hidden), linear_hidden_shapes=hidden,
                        num_feature_extractor_layers=1, feature_extractor_latent_size=8, use_bias=True,
                        use_batch_norm=False, feature_extractor="LSTM", last_layer_std=0.01),
        F.DynamicConfig(0, 0, 0, 0),
        processors=1, device="cpu", experiment_path=tmpdir, verbose=False, central_critic=True,
        central_actor=True, normalize_rewards=spec.get("normalize_rewards", False), normalize_actions=True,
        normalize_observations=True, sequence_wise_normalization=True, dtype=torch.float32,
        render_size=[8, 8])


class _Helper:
    """Stands in for EnvironmentHelper: the hot path only reads `.run` from it."""

    def __init__(self, run):
        self.run = run


def run_case(name: str):
    from oracle import shims
    shims.install()
    from oracle.ppo_oracle import synthetic_rollout
    from tensordict import TensorDict  # the shim
    spec = CASES[name]
    torch.manual_seed(spec["seed"])
    torch.set_num_threads(1)
    tmpdir = tempfile.mkdtemp(prefix="golden_")
    run = _build_run(spec, tmpdir)
    from utils.logger import Logger
    Logger.log("golden", episode=0, log_type=Logger.TRAINING_TYPE, path=tmpdir)
    from entities.algorithms.ppo import PPO

    N, T = spec["N"], spec["T"]
    D, A = spec.get("D", 4), spec.get("A", 2)
    roll = synthetic_rollout(N, T, D, A, seed=1000 + spec["seed"], p_term=0.05,
                             reward_f64=spec.get("reward_f64", False))
    roll["current_state"] = roll["current_state"].reshape(N, T, D)  # window_length 1, flattened
    out = {}
    if spec["kind"] == "obsnorm":
        # EnvironmentHelper.normalize_state / _normalize run verbatim (running_gym_sequential_vectorized.py:61-82) on a
        # stand-in `self`; the cast + permute of get_state (:89-91) follow, since get_state itself needs a live env.
        import types
        from environments.humanoid.running_gym_sequential_vectorized import EnvironmentHelper as RefHelper
        g = torch.Generator().manual_seed(spec["seed"])
        obs = torch.randn(spec["N"], 376, 5, generator=g, dtype=torch.float64) * 3.0 + 0.5
        obs[2, 253:270, 1] = 0.25  # a constant segment: std == 0 -> 1
        stand_in = types.SimpleNamespace()
        stand_in._normalize = lambda x: RefHelper._normalize(stand_in, x)
        normed = RefHelper.normalize_state(stand_in, obs.clone())
        out["in_observation"] = obs.numpy()
        out["state"] = normed.to(torch.float32).permute(0, 2, 1).contiguous().numpy()
    elif spec["kind"] == "sac":
        # SoftActorCritic.train runs verbatim (soft_actor_critic.py:33-118) with the agent's two module-level names
        # re-pointed at the MLP classes models.linear.actor.Actor / models.linear.q_network.QNetwork (as committed it binds
        # the Transformer variants).  torch.randperm and the standard-normal draws inside Normal.rsample are recorded.
        import torch.distributions.normal as normal_mod
        import entities.agents.soft_actor_critic_agent as sac_agent_mod
        from models.linear.actor import Actor as LinearActor
        from models.linear.q_network import QNetwork as LinearQ
        from oracle.sac_oracle import synthetic_replay
        sac_agent_mod.Actor = LinearActor
        sac_agent_mod.QNetwork = LinearQ
        agent = sac_agent_mod.SoftActorCriticAgent()
        from entities.algorithms.soft_actor_critic import SoftActorCritic
        algo = SoftActorCritic(_Helper(run), agent)  # hard_update(target, online) happens here
        for k, v in agent.networks.state_dict().items():
            out["init/" + k] = v.detach().clone().numpy()
        W = spec["W"]
        replay = synthetic_replay(N, T, (W, D), A, seed=2000 + spec["seed"])
        for k, v in replay.items():
            out["mem/" + k] = v.numpy()
        perms, eps = [], []
        real_randperm, real_std_normal = torch.randperm, normal_mod._standard_normal

        def recording_randperm(*a, **k):
            p = real_randperm(*a, **k)
            perms.append(p.clone())
            return p

        def recording_std_normal(*a, **k):
            e = real_std_normal(*a, **k)
            eps.append(e.clone())
            return e

        torch.randperm = recording_randperm
        normal_mod._standard_normal = recording_std_normal
        losses = []
        try:
            for call in range(spec["calls"]):
                mem = TensorDict({k: v.clone() for k, v in replay.items()}, batch_size=(N, T))
                losses.append(list(algo.train(mem, call)))
        finally:
            torch.randperm = real_randperm
            normal_mod._standard_normal = real_std_normal
        out["perms"] = torch.stack(perms).numpy()
        out["eps"] = torch.stack(eps).numpy()            # [2 * calls, B, A]: (next-state draw, policy draw) per call
        out["losses"] = np.array(losses, dtype=np.float64)
        for k, v in agent.networks.state_dict().items():
            out["final/" + k] = v.detach().clone().numpy()
        for oname in ("actor", "online_critic"):
            sd = agent.optimizers[oname].state_dict()
            for pid, st in sd["state"].items():
                out[f"opt/{oname}/{pid}/exp_avg"] = st["exp_avg"].numpy()
                out[f"opt/{oname}/{pid}/exp_avg_sq"] = st["exp_avg_sq"].numpy()
        out["cfg"] = np.array([spec["B"], spec.get("lr", 1e-4), 0.99, 0.05, 0.005, 1, 1.0, spec.get("out_max", 1.0), spec["W"]],
                              dtype=np.float64)  # batch, lr, gamma, alpha, tau, target_update_interval, max_grad_norm, out_max, window
        out["hidden"] = np.array(spec["hidden"], dtype=np.int64)
        out["activation"] = np.array(spec["act"].lower())
    elif spec["kind"] == "adv":
        mem = TensorDict({k: v.clone() for k, v in roll.items()}, batch_size=(N, T))
        algo = PPO(_Helper(run), agent=None)
        algo.calculate_advantages(mem)
        for k in ("reward", "current_state_value", "next_state_value", "terminated"):
            out["in_" + k] = roll[k].numpy()
        out["advantage"] = mem["advantage"].numpy()
        out["value_target"] = mem["current_state_value_target"].numpy()
        out["cfg"] = np.array([0.99, 0.98, float(spec.get("normalize_rewards", False)),
                               float(spec.get("normalize_advantage", False)),
                               spec.get("advantage_scaler", 1.0)], dtype=np.float64)
    else:
        import entities.agents.ppo_agent as ppo_agent_mod
        from models.linear.actor import Actor as LinearActor
        from models.critic import Critic as MLPCritic
        ppo_agent_mod.Actor = LinearActor
        ppo_agent_mod.Critic = MLPCritic
        agent = ppo_agent_mod.PPOAgent()
        for k, v in agent.networks.state_dict().items():
            out["init/" + k] = v.detach().clone().numpy()
        algo = PPO(_Helper(run), agent)
        # old log-prob under the initial policy
        with torch.no_grad():
            mean, std = agent.networks["actor"](roll["current_state"].reshape(N * T, D))
            logp = torch.distributions.Normal(mean, std).log_prob(roll["action"].reshape(N * T, A)).sum(dim=1)
        roll["action_log_prob"] = (logp + 0.05 * torch.randn(N * T)).reshape(N, T)
        mem = TensorDict({k: v.clone() for k, v in roll.items()}, batch_size=(N, T))
        algo.calculate_advantages(mem)
        perms = []
        real_randperm = torch.randperm

        def recording_randperm(*a, **k):
            p = real_randperm(*a, **k)
            perms.append(p.clone())
            return p

        messages = []
        real_log = Logger.log

        def recording_log(message, *a, **k):
            messages.append(message)
            k["print_message"] = False
            return real_log(message, *a, **k)

        torch.randperm = recording_randperm
        Logger.log = staticmethod(recording_log)
        try:
            algo.train(mem)
        finally:
            torch.randperm = real_randperm
            Logger.log = staticmethod(real_log)
        for k in ("current_state", "action", "action_log_prob", "reward", "current_state_value",
                  "next_state_value", "terminated", "advantage", "current_state_value_target"):
            out["mem/" + k] = mem[k].numpy()
        out["perms"] = torch.stack(perms).numpy()
        for k, v in agent.networks.state_dict().items():
            out["final/" + k] = v.detach().clone().numpy()
        for oname, opt in agent.optimizers.items():
            sd = opt.state_dict()
            out[f"opt/{oname}/lr"] = np.array(sd["param_groups"][0]["lr"], dtype=np.float64)
            for pid, st in sd["state"].items():
                out[f"opt/{oname}/{pid}/exp_avg"] = st["exp_avg"].numpy()
                out[f"opt/{oname}/{pid}/exp_avg_sq"] = st["exp_avg_sq"].numpy()
                out[f"opt/{oname}/{pid}/step"] = np.array(float(st["step"]), dtype=np.float64)
        msg = [m for m in messages if m.startswith("Actor Loss")][-1].split()
        out["logged_losses"] = np.array([float(msg[2]), float(msg[5])], dtype=np.float64)
        out["cfg"] = np.array([spec["B"], spec["epochs"], spec.get("lr", 1e-4), spec.get("clip", 0.1),
                               spec.get("ent", 1e-4), spec.get("out_max", 1.0)], dtype=np.float64)
        out["hidden"] = np.array(spec["hidden"], dtype=np.int64)
        out["activation"] = np.array(spec["act"].lower())
    os.makedirs(GOLDEN_DIR, exist_ok=True)
    path = os.path.join(GOLDEN_DIR, f"ref_{name}.npz")
    np.savez_compressed(path, **out)
    print(f"wrote {path}: {len(out)} arrays, {os.path.getsize(path)} bytes")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--case")
    args = ap.parse_args()
    if args.case:
        run_case(args.case)
        return
    root = os.path.dirname(GOLDEN_DIR.rstrip("/")).rsplit("/tests", 1)[0]
    for name in CASES:
        subprocess.check_call([sys.executable, "-m", "oracle.make_golden", "--case", name], cwd=root)


if __name__ == "__main__":
    main()
