"""Launch ONE hot-path kernel a few times at its roofline size (the command ncu wraps; see profiles/README.md).

    python profiles/kernels.py gae|gae_norm|gather|adam|update [--precision fp32|bf16]
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mujoco_reinforcement_learning_b200 as pkg  # noqa: E402

which = sys.argv[1]
precision = sys.argv[sys.argv.index("--precision") + 1] if "--precision" in sys.argv else "fp32"
batch = int(sys.argv[sys.argv.index("--batch") + 1]) if "--batch" in sys.argv else 4096
dev = "cuda"
torch.manual_seed(0)
if which in ("gae", "gae_norm"):
    n, t = 65536, 1024
    r, v, vn = (torch.randn(n, t, 1, device=dev) for _ in range(3))
    term = torch.rand(n, t, device=dev) < 0.01
    for _ in range(3):
        pkg.calculate_advantages(r, v, vn, term, 0.99, 0.98, normalize_advantage=(which == "gae_norm"))
elif which == "gather":
    m = 4096 * 128
    obs, act, s = torch.randn(m, 376, device=dev), torch.randn(m, 17, device=dev), torch.randn(m, device=dev)
    idx = torch.randperm(m, device=dev)
    for _ in range(3):
        pkg.gather_minibatch(idx, obs, act, s, s, s, check=False)
elif which == "adam":
    n = 64 * 1024 * 1024
    p, g, m_, v_ = (torch.randn(n, device=dev) for _ in range(4))
    v_.abs_()
    for i in range(3):
        pkg.adam_step_(p, g, m_, v_, i + 1, 1e-4)
elif which == "update":
    B, N, T = batch, max(512, batch // 64 * 2), 64
    run = pkg.Run(training_config=pkg.TrainingConfig(batch_size=B, epochs_per_iteration=1),
                  environment_config=pkg.EnvironmentConfig(maximum_timesteps=T, num_envs=N),
                  network_config=pkg.NetworkConfig(input_shape=376, output_shape=17, linear_hidden_shapes=[256, 256], critic_hidden_shapes=[256, 256]),
                  gemm_precision=precision)
    agent = pkg.PPOAgent(run, max_batch=B)
    M = N * T
    mem = pkg.RolloutMemory({"current_state": torch.randn(N, T, 376, device=dev), "action": torch.randn(N, T, 17, device=dev),
                             "action_log_prob": torch.randn(N, T, device=dev) - 20, "advantage": torch.randn(N, T, 1, device=dev),
                             "current_state_value_target": torch.randn(N, T, 1, device=dev)}, (N, T))
    algo = pkg.PPO(type("H", (), {"run": run})(), agent)
    algo.train(mem, perms=torch.randperm(M)[None])
torch.cuda.synchronize()
print("done", which)
