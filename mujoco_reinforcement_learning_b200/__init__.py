"""B200-native PPO update path: drop-in surface of aminrezaee/mujoco_reinforcement_learning's PPO
(`Actor`, `Critic`, `PPOAgent`, `PPO`, `generalized_advantage_estimate`) over hand-written sm_100a kernels."""
from .config import (DynamicConfig, EnvironmentConfig, NetworkConfig, PPOConfig, Run, SACConfig, TrainingConfig)  # noqa: F401
from .functional import (adam_step_, calculate_advantages, gather_minibatch, gather_rows,  # noqa: F401
                         generalized_advantage_estimate, normalize_state, polyak_update_)
from .memory import RolloutMemory  # noqa: F401
from .models import Actor, ActorCriticEngine, Critic, NetworkBlock, create_network  # noqa: F401
from .agent import FusedAdam, PPOAgent  # noqa: F401
from .ppo import PPO  # noqa: F401
from .sac import QNetwork, SoftActorCritic, SoftActorCriticAgent  # noqa: F401

__version__ = "0.1.0"
