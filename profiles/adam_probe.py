"""Per-launch time of the optimizer kernel at the real parameter count (329 k) as a function of the number of split-K partials
it sums, back to back on one stream (CUDA events around 200 launches).  python profiles/adam_probe.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mujoco_reinforcement_learning_b200 import _lib  # noqa: E402

lib = _lib.load()
n = 329280
dev = "cuda"
p, m, v = torch.randn(n, device=dev), torch.zeros(n, device=dev), torch.ones(n, device=dev)
for parts in (1, 2, 5, 9, 18):
    g = torch.randn(parts, n, device=dev) * 1e-3
    def run(k):
        for i in range(k):
            _lib.check(lib.b200ppo_adam_step(_lib.ptr(p), _lib.ptr(g), parts, n, _lib.ptr(m), _lib.ptr(v), n, 1e-4, 0.9, 0.999, 1e-8, i + 1,
                                             _lib.stream_ptr()), "adam")
    run(20)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); run(200); b.record()
    torch.cuda.synchronize()
    print(f"{parts:2d} partials: {a.elapsed_time(b) / 200 * 1e3:7.2f} us per launch (329 k parameters, L2-resident)")
