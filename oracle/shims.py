"""`sys.modules` shims that let the reference be imported VERBATIM — TEST INFRASTRUCTURE ONLY.

Build-container only (needs `/root/reference`).  The reference imports `tensordict`,
`torchrl`, `mlflow`, `mediapy` and `gymnasium`, none of which is installed or
installable here.  This module registers the smallest stand-ins that make
`entities.algorithms.ppo.PPO.calculate_advantages` / `.train` and the MLP modules run
unmodified:

 * `tensordict.TensorDict` — dict of tensors sharing leading batch dims, with exactly the
   operations ppo.py and soft_actor_critic.py touch (`[]` by key / index tensor / slice, `[]=`, `len`,
   `view(-1)`, `torch.clone`).
 * `torchrl.objectives.value.functional.generalized_advantage_estimate` — the restatement in
   `oracle/ppo_oracle.py` (torchrl's source is not available: parity unpinned for it).
 * `mlflow`, `mediapy`, `gymnasium(.core)` — import-time placeholders, never called on the path.
"""
from __future__ import annotations

import sys
import types

import torch

REFERENCE_SRC = "/root/reference/src"


class TensorDict:
    def __init__(self, source, batch_size=None, **_):
        self._d = dict(source)
        if batch_size is None:
            batch_size = ()
        self.batch_size = torch.Size(batch_size)

    def keys(self):
        return self._d.keys()

    def __len__(self):
        return self.batch_size[0]

    def __getitem__(self, key):
        if isinstance(key, str):
            return self._d[key]
        out = {k: v[key] for k, v in self._d.items()}
        probe = torch.empty(self.batch_size, device="meta")[key]
        return TensorDict(out, batch_size=probe.shape)

    def __setitem__(self, key, value):
        assert isinstance(key, str)
        assert value.shape[:len(self.batch_size)] == self.batch_size, (key, value.shape, self.batch_size)
        self._d[key] = value

    @classmethod
    def __torch_function__(cls, func, types, args=(), kwargs=None):
        # `torch.clone(memory)` — soft_actor_critic.py:40
        if func is torch.clone:
            td = args[0]
            return TensorDict({k: v.clone() for k, v in td._d.items()}, batch_size=td.batch_size)
        return NotImplemented

    def view(self, *shape):
        assert shape == (-1,)
        nb = len(self.batch_size)
        out = {k: v.reshape(-1, *v.shape[nb:]) for k, v in self._d.items()}
        return TensorDict(out, batch_size=(self.batch_size.numel(),))


def install(src: str = REFERENCE_SRC):
    """Register the shims and put the reference's `src/` (or its staged copy `oracle/_ref/src`) on sys.path."""
    from oracle.ppo_oracle import generalized_advantage_estimate

    def mod(name, **attrs):
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        sys.modules[name] = m
        return m

    mod("tensordict", TensorDict=TensorDict)
    mod("torchrl")
    mod("torchrl.objectives")
    mod("torchrl.objectives.value")
    mod("torchrl.objectives.value.functional", generalized_advantage_estimate=generalized_advantage_estimate)
    noop = lambda *a, **k: None
    mod("mlflow", log_metric=noop, log_metrics=noop, log_artifact=noop, set_tag=noop, set_tags=noop)
    mod("mediapy", write_video=noop, show_video=noop)
    core = mod("gymnasium.core", Env=object)
    mod("gymnasium", core=core, Env=object)
    if src not in sys.path:
        sys.path.insert(0, src)
