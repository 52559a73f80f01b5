"""Phase timeline of one fp32-tolerance GEMM launch (B200PPO_SPLIT_TRACE=<n-th launch>): the first-layer forward shape."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mujoco_reinforcement_learning_b200 import _lib
lib = _lib.load()
M, N, K = 32768, 256, 376
A = torch.randn(M, K, device="cuda"); B = torch.randn(N, K, device="cuda"); C = torch.empty(M, N, device="cuda")
for i in range(3):
    _lib.check(lib.b200ppo_debug_gemm_split(_lib.ptr(A), _lib.ptr(B), _lib.ptr(C), None, M, N, K, 0, 0, 1, _lib.stream_ptr()), "x")
torch.cuda.synchronize()
