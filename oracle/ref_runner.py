"""Runs the reference's OWN `PPO.calculate_advantages` / `PPO.train` (ppo.py:62-154) from `oracle/_ref` — TEST
INFRASTRUCTURE ONLY (the timed CPU arm of `bench.py`; see `oracle/build_ref.py`).

What executes unmodified from the staged sources: `entities.algorithms.ppo.PPO`, `entities.agents.ppo_agent.PPOAgent`
(with its two module-level names `Actor` / `Critic` re-pointed at the reference's MLP classes
`models.linear.actor.Actor` and `models.critic.Critic`, as `oracle/make_golden.py` does: as committed the agent binds
the LSTM variants), `models.network_block_creator`, `entities.features.Run`.  Third-party imports the
image lacks are shimmed (`oracle/shims.py`); the torchrl GAE function is the restatement in `oracle/ppo_oracle.py`.
Note: the reference's `Critic` hard-codes 128x128 hidden layers (models/critic.py:13-14) whatever the config says.
"""
from __future__ import annotations

import tempfile

import torch

from oracle import build_ref, shims


class ReferencePPO:
    """The reference's PPO + PPOAgent for an MLP actor `obs_dim -> hidden -> act_dim` (one per process: `Run` is a
    process-wide singleton, utils/type_utils.py:1-7)."""

    def __init__(self, obs_dim, act_dim, hidden, batch_size, epochs, n_envs, steps, activation="Tanh"):
        if not build_ref.available():
            raise RuntimeError("oracle/_ref is not staged (python -m oracle.build_ref in the build container)")
        shims.install(build_ref.REF_SRC)
        from entities import features as F
        self.tmpdir = tempfile.mkdtemp(prefix="ref_arm_")
        self.run = F.Run(
            F.RewardConfig(),
            F.TrainingConfig(iteration_count=1, learning_rate=1e-4, weight_decay=1e-4, batch_size=batch_size,
                             epochs_per_iteration=epochs, minimum_learning_rate=1e-4),
            F.PPOConfig(max_grad_norm=1.0, clip_epsilon=0.1, gamma=0.99, lmbda=0.98, entropy_eps=1e-4, advantage_scaler=1.0,
                        normalize_advantage=False, critic_coeffiecient=1.0),
            F.SACConfig(1.0, 0.99, 0.05, 0.005, 999, 1, False),
            F.EnvironmentConfig(maximum_timesteps=steps, num_envs=n_envs, window_length=1),
            F.AgentConfig(sub_action_count=1),
            F.NetworkConfig(input_shape=obs_dim, output_shape=act_dim, output_max_value=1.0,
                            activation_class=getattr(torch.nn, activation), num_linear_layers=len(hidden),
                            linear_hidden_shapes=list(hidden), num_feature_extractor_layers=1,
                            feature_extractor_latent_size=8, use_bias=True, use_batch_norm=False, feature_extractor="LSTM",
                            last_layer_std=0.01),
            F.DynamicConfig(0, 0, 0, 0),
            processors=1, device="cpu", experiment_path=self.tmpdir, verbose=False, central_critic=True, central_actor=True,
            normalize_rewards=False, normalize_actions=True, normalize_observations=True, sequence_wise_normalization=True,
            dtype=torch.float32, render_size=[8, 8])
        from utils.logger import Logger
        Logger.log("reference arm", episode=0, log_type=Logger.TRAINING_TYPE, path=self.tmpdir)
        real_log = Logger.log

        def quiet_log(message, *a, **k):  # ppo.py:149-153 prints its loss line; keep stdout to bench.py's one JSON line
            k["print_message"] = False
            return real_log(message, *a, **k)

        Logger.log = staticmethod(quiet_log)
        import entities.agents.ppo_agent as ppo_agent_mod
        from models.critic import Critic as MLPCritic
        from models.linear.actor import Actor as LinearActor
        ppo_agent_mod.Actor = LinearActor
        ppo_agent_mod.Critic = MLPCritic
        self.agent = ppo_agent_mod.PPOAgent()
        from entities.algorithms.ppo import PPO
        helper = type("Helper", (), {"run": self.run})()
        self.algo = PPO(helper, self.agent)
        from tensordict import TensorDict  # the shim
        self.TensorDict = TensorDict
        self.n_envs, self.steps = n_envs, steps

    def memory(self, roll):
        """The `[N, T]` container `PPO.rollout` would return (ppo.py:60), over the given leaves."""
        return self.TensorDict({k: v for k, v in roll.items()}, batch_size=(self.n_envs, self.steps))

    def calculate_advantages(self, mem):
        self.algo.calculate_advantages(mem)

    def train(self, mem, epochs=None):
        if epochs is not None:
            self.run.training_config.epochs_per_iteration = epochs
        self.algo.train(mem)
