"""CPU-side checks of the boundary: the C-ABI library loads and exports every symbol the header declares,
the ctypes table matches the header, inputs on the wrong device are rejected (no fallback), module init
matches the reference's RNG stream, RolloutMemory semantics."""
import os
import re

import numpy as np
import pytest
import torch

import mujoco_reinforcement_learning_b200 as pkg
from mujoco_reinforcement_learning_b200 import _lib
from tests._util import load_golden, sub

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_functions():
    text = open(os.path.join(ROOT, "include", "b200ppo.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(b200ppo_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    lib = _lib.load()
    names = _header_functions()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/b200ppo.h but not exported"
    assert lib.b200ppo_version() == 100


def test_ctypes_table_matches_header():
    assert sorted(_lib.SIGNATURES) == _header_functions()


def test_cpu_tensors_are_rejected_not_routed_elsewhere():
    x = torch.zeros(2, 8, 1)
    b = torch.zeros(2, 8, 1, dtype=torch.bool)
    with pytest.raises(RuntimeError, match="CUDA"):
        pkg.generalized_advantage_estimate(0.99, 0.95, x, x, x, b, b)
    with pytest.raises(RuntimeError, match="CUDA"):
        pkg.gather_rows(x, torch.zeros(1, dtype=torch.int64))
    with pytest.raises(RuntimeError, match="CUDA"):
        pkg.adam_step_(x.view(-1), x.view(-1), x.view(-1), x.view(-1), 1, 1e-3)


def test_gae_shape_error_matches_torchrl():
    x = torch.zeros(2, 8, 1)
    b = torch.zeros(2, 8, 1, dtype=torch.bool)
    with pytest.raises(RuntimeError, match="unique shape"):
        pkg.generalized_advantage_estimate(0.99, 0.95, x, x, x[:, :4], b, b)


@pytest.mark.skipif(torch.cuda.is_available(), reason="only meaningful without a GPU")
def test_engine_fails_loudly_without_gpu():
    run = pkg.Run(network_config=pkg.NetworkConfig(input_shape=4, output_shape=2, linear_hidden_shapes=[8, 8]))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        pkg.PPOAgent(run)


@pytest.mark.parametrize("name,seed", [("train_tanh64", 21), ("train_relu3", 22), ("train_tanh96", 23)])
def test_module_init_matches_reference_rng_stream(name, seed):
    """Same seed -> same initial weights as the reference's Actor()/Critic() (construction order and draws)."""
    g = load_golden(name)
    init = sub(g, "init/")
    D = init["actor.actor.first_layers.0.weight"].shape[1]
    A = init["actor.actor_logstd"].shape[0]
    act = {"tanh": torch.nn.Tanh, "relu": torch.nn.ReLU}[str(g["activation"])]
    hidden = [int(h) for h in g["hidden"]]
    run = pkg.Run(network_config=pkg.NetworkConfig(input_shape=D, output_shape=A, activation_class=act,
                                                   num_linear_layers=len(hidden), linear_hidden_shapes=hidden,
                                                   critic_hidden_shapes=[128, 128]))
    torch.manual_seed(seed)
    nets = torch.nn.ModuleDict({"actor": pkg.Actor(run), "critic": pkg.Critic(run)})
    sd = nets.state_dict()
    assert sorted(sd.keys()) == sorted(init.keys())
    for k, v in sd.items():
        # same draws; the QR inside orthogonal_ may round differently with another BLAS thread count
        np.testing.assert_allclose(v.numpy(), init[k], rtol=0, atol=1e-5, err_msg=k)


def test_rollout_memory_semantics():
    m = pkg.RolloutMemory({"a": torch.arange(24.).reshape(2, 3, 4), "b": torch.zeros(2, 3, dtype=torch.bool)}, (2, 3))
    assert len(m) == 2
    f = m.view(-1)
    assert len(f) == 6 and f["a"].shape == (6, 4) and f["b"].shape == (6,)
    assert torch.equal(f["a"][4], m["a"][1, 1])  # flat index n*T + t
    s = f[1:4]
    assert len(s) == 3 and torch.equal(s["a"], f["a"][1:4])
    with pytest.raises(RuntimeError):
        m["c"] = torch.zeros(3, 2)
    c = pkg.RolloutMemory.cat([pkg.RolloutMemory({"x": torch.ones(2, 1, 5) * i}, (2, 1)) for i in range(3)], dim=1)
    assert c.batch_size == (2, 3) and c["x"][0, 2, 0] == 2
