"""Times the normalised and the plain GAE scan at 65536 x 1024 (1.41 GB algorithmic at 21 B per element)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mujoco_reinforcement_learning_b200 as pkg
n, t = 65536, 1024
dev = "cuda"
r, v, vn = (torch.randn(n, t, 1, device=dev) for _ in range(3))
term = torch.rand(n, t, device=dev) < 0.01
for norm in (False, True):
    ts = []
    for i in range(8):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); pkg.calculate_advantages(r, v, vn, term, 0.99, 0.98, normalize_advantage=norm); b.record()
        torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    ts.sort(); med = ts[len(ts) // 2]
    print(f"normalize_advantage={norm}: {med:.4f} ms  {21 * n * t / 1e9 / (med * 1e-3):.0f} GB/s")
