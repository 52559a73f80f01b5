"""Configuration objects with the reference's field names (src/entities/features.py:17-122).

Only the fields the PPO hot path reads are kept; `Run.instance()` returns the most recently constructed
`Run`, which is how the reference's code reaches its process-wide singleton (features.py:129-133).
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import List, Optional

import torch


@dataclass
class TrainingConfig:
    learning_rate: float = 1e-4
    batch_size: int = 500
    epochs_per_iteration: int = 10
    iteration_count: int = 3000
    weight_decay: float = 1e-4          # dead field in the reference (never read on the PPO path)
    minimum_learning_rate: float = 1e-4  # dead field


@dataclass
class PPOConfig:
    max_grad_norm: float = 1.0   # read at ppo.py:136-137, after both optimiser steps: no effect on parameters
    clip_epsilon: float = 0.1
    gamma: float = 0.99
    lmbda: float = 0.98
    entropy_eps: float = 1e-4
    advantage_scaler: float = 1.0
    normalize_advantage: bool = False
    critic_coeffiecient: float = 1.0  # dead field (spelling as in the reference)


@dataclass
class SACConfig:
    """src/entities/features.py:91-98."""
    max_grad_norm: float = 1.0   # dead field: SoftActorCritic.train clips with ppo_config.max_grad_norm (:67, :84)
    gamma: float = 0.99
    alpha: float = 0.05
    tau: float = 0.005
    memory_capacity: int = 999
    target_update_interval: int = 1
    automatic_entropy_tuning: bool = False


@dataclass
class EnvironmentConfig:
    maximum_timesteps: int = 500
    num_envs: int = 5
    window_length: int = 1


@dataclass
class NetworkConfig:
    input_shape: int = 348
    output_shape: int = 17
    output_max_value: float = 1.0
    activation_class: type = torch.nn.Tanh
    num_linear_layers: int = 2
    linear_hidden_shapes: List[int] = field(default_factory=lambda: [256, 256])
    critic_hidden_shapes: Optional[List[int]] = None  # None = [128, 128], what the reference hard-codes (models/critic.py:13-14)
    use_bias: bool = True
    use_batch_norm: bool = False
    last_layer_std: float = 0.01


@dataclass
class DynamicConfig:
    current_episode: int = 0
    current_episode_timestep: int = 0
    current_timestep: int = 0
    best_reward: float = 0.0

    def next_episode(self):
        self.current_episode = int(self.current_episode + 1)


@dataclass
class Run:
    training_config: TrainingConfig = field(default_factory=TrainingConfig)
    ppo_config: PPOConfig = field(default_factory=PPOConfig)
    sac_config: SACConfig = field(default_factory=SACConfig)
    environment_config: EnvironmentConfig = field(default_factory=EnvironmentConfig)
    network_config: NetworkConfig = field(default_factory=NetworkConfig)
    dynamic_config: DynamicConfig = field(default_factory=DynamicConfig)
    device: str = "cuda"
    normalize_rewards: bool = False
    dtype: torch.dtype = torch.float32
    experiment_path: str = ""
    gemm_precision: str = "fp32"  # "fp32" (1e-5 parity) or "bf16" (tcgen05 tensor cores, 2e-2)

    _instance = None

    def __post_init__(self):
        Run._instance = self

    @staticmethod
    def instance() -> "Run":
        return Run._instance
