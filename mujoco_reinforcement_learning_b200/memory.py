"""Rollout container with the few TensorDict operations the reference's PPO uses
(tensordict 0.6.2 at src/entities/algorithms/ppo.py:49-50,60,90-91,99,104,106)."""
from __future__ import annotations

from typing import Dict, Iterable

import torch

from . import functional as F


class RolloutMemory:
    """Dict of tensors sharing leading batch dims ([N_envs, T] for a rollout)."""

    def __init__(self, source: Dict[str, torch.Tensor], batch_size):
        self._d = dict(source)
        self.batch_size = torch.Size(batch_size)
        for k, v in self._d.items():
            if tuple(v.shape[:len(self.batch_size)]) != tuple(self.batch_size):
                raise RuntimeError(f"leaf {k!r} has shape {tuple(v.shape)}, batch dims {tuple(self.batch_size)} expected")

    def keys(self) -> Iterable[str]:
        return self._d.keys()

    def __contains__(self, key):
        return key in self._d

    def __len__(self):
        return self.batch_size[0]

    def __getitem__(self, key):
        if isinstance(key, str):
            return self._d[key]
        if isinstance(key, slice):
            out = {k: v[key] for k, v in self._d.items()}
            n = len(range(*key.indices(self.batch_size[0])))
            return RolloutMemory(out, (n, *self.batch_size[1:]))
        if isinstance(key, torch.Tensor) and key.dtype == torch.int64 and key.dim() == 1:
            # memory[idx]: gather every leaf along dim 0 (ppo.py:104) with the CUDA row-gather
            idx = key.to(next(iter(self._d.values())).device)
            out = {k: F.gather_rows(v, idx) for k, v in self._d.items()}
            return RolloutMemory(out, (idx.numel(), *self.batch_size[1:]))
        raise TypeError(f"unsupported index {type(key)}")

    def __setitem__(self, key: str, value: torch.Tensor):
        if tuple(value.shape[:len(self.batch_size)]) != tuple(self.batch_size):
            raise RuntimeError(f"leaf {key!r} has shape {tuple(value.shape)}, batch dims {tuple(self.batch_size)} expected")
        self._d[key] = value

    def view(self, *shape):
        if shape != (-1,):
            raise NotImplementedError("only view(-1) is used by the PPO path")
        nb = len(self.batch_size)
        out = {k: v.reshape(-1, *v.shape[nb:]) for k, v in self._d.items()}
        return RolloutMemory(out, (self.batch_size.numel(),))

    @staticmethod
    def cat(items, dim: int = 1) -> "RolloutMemory":
        """torch.cat(memory, dim=1) of per-step items (ppo.py:60)."""
        keys = list(items[0].keys())
        out = {k: torch.cat([it[k] for it in items], dim=dim) for k in keys}
        bs = list(items[0].batch_size)
        bs[dim] = sum(it.batch_size[dim] for it in items)
        return RolloutMemory(out, bs)
