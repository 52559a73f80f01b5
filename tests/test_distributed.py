"""N > 1 path: host-side sharding logic on CPU (gloo, world_size 2) and, when >= 2 GPUs are visible, the real
NCCL data-parallel update against the single-GPU result."""
import os
import subprocess
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from mujoco_reinforcement_learning_b200 import distributed as D

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_rank_rows_partition_every_minibatch():
    for world in (1, 2, 4, 8):
        gb, nb = 64, 5
        for i in range(nb):
            rows = []
            for r in range(world):
                rows += list(D.rank_rows(i, gb, world, r))
            assert rows == list(range(i * gb, (i + 1) * gb))  # union over ranks == the reference's minibatch slice


def _gloo_worker(rank, world, port, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sys.path.insert(0, ROOT)
    from oracle import ppo_oracle as O  # the CPU stand-in for the per-rank compute in this host-logic test
    torch.manual_seed(0)
    n_local, T, Dm, A, gb = 4, 8, 5, 2, 16
    N, M = n_local * world, n_local * world * T
    g = torch.Generator().manual_seed(3)
    full = {"current_state": torch.randn(N * T, Dm, generator=g), "action": torch.randn(N * T, A, generator=g),
            "action_log_prob": torch.randn(N * T, generator=g) * 0.1 - 2.5, "advantage": torch.randn(N * T, 1, generator=g),
            "current_state_value_target": torch.randn(N * T, 1, generator=g)}
    perm = torch.randperm(M, generator=g)
    # (1) all-gather of env slabs reproduces the global env-major buffer
    mine = {k: v[rank * n_local * T:(rank + 1) * n_local * T].clone() for k, v in full.items()}
    gathered = D.all_gather_fields(mine)
    ok = all(torch.equal(gathered[k], full[k]) for k in full)
    # (2) per-rank gradients with the GLOBAL divisor, summed over ranks == the full-minibatch gradient
    cfg = O.OracleConfig(obs_dim=Dm, act_dim=A, actor_hidden=[8, 8], critic_hidden=[8, 8], batch_size=gb)
    agent = O.OracleAgent(cfg)
    idx_full = perm[:gb]
    rows = list(D.rank_rows(0, gb, world, rank))
    idx = perm[rows]
    b = {k: v[idx] for k, v in gathered.items()}
    _, _, grads, _, _ = O.minibatch_grads(agent, b["current_state"], b["action"], b["action_log_prob"], b["advantage"],
                                          b["current_state_value_target"])
    lb = gb // world
    ent_fix = {}
    for k in grads:  # oracle losses are means over the LOCAL rows: rescale to the global divisor
        grads[k] = grads[k] * (lb / gb)
    # the entropy term is batch-independent: after rescaling every rank carries 1/world of it, like rank_share
    flat = torch.cat([grads[k].reshape(-1) for k in sorted(grads)])
    dist.all_reduce(flat)
    bf = {k: v[idx_full] for k, v in gathered.items()}
    _, _, gfull, _, _ = O.minibatch_grads(agent, bf["current_state"], bf["action"], bf["action_log_prob"], bf["advantage"],
                                          bf["current_state_value_target"])
    ref = torch.cat([gfull[k].reshape(-1) for k in sorted(gfull)])
    err = ((flat - ref).abs().max() / ref.abs().max()).item()
    ret[rank] = (ok, err)
    dist.destroy_process_group()


def test_slice_perms_for_rank_is_rank_rows_of_every_minibatch():
    """The rank-sliced permutation layout of `b200ppo_set_perm_layout(ctx, 1)`: slot (e, i, j) of rank r is global slot
    i * batch + r * (batch / world) + j of epoch e; the tail beyond floor(M / batch) minibatches is left behind."""
    M, GB, world, E = 1000, 96, 4, 3
    g = torch.Generator().manual_seed(3)
    perms = torch.stack([torch.randperm(M, generator=g) for _ in range(E)])
    seen = [[] for _ in range(E)]
    for r in range(world):
        sl = D.slice_perms_for_rank(perms, GB, world, r)
        assert sl.shape == (E, M // GB, GB // world) and sl.dtype == torch.int64
        for e in range(E):
            for i in range(M // GB):
                assert torch.equal(sl[e, i], perms[e][list(D.rank_rows(i, GB, world, r))])
            seen[e].append(sl[e].reshape(-1))
    for e in range(E):  # the ranks' slices together are exactly the kept part of the permutation
        assert torch.equal(torch.cat(seen[e]).sort().values, perms[e][:(M // GB) * GB].sort().values)


def test_bind_host_to_gpu_never_fails_and_stays_inside_the_allowed_cpus():
    """The NUMA binding is an optimisation: without a GPU / NVML it reports False and leaves the affinity alone;
    with one it may only narrow the allowed set."""
    before = os.sched_getaffinity(0)
    try:
        ok = D.bind_host_to_gpu("cuda:0")
        after = os.sched_getaffinity(0)
        assert isinstance(ok, bool)
        assert after <= before and len(after) >= 1
        if not ok:
            assert after == before
    finally:
        os.sched_setaffinity(0, before)


def test_gloo_world2_allgather_and_gradient_sum():
    world = 2
    mgr = mp.Manager()
    ret = mgr.dict()
    port = 29500 + (os.getpid() % 400)
    mp.spawn(_gloo_worker, args=(world, port, ret), nprocs=world, join=True)
    for r in range(world):
        ok, err = ret[r]
        assert ok, "all-gather did not reproduce the env-major buffer"
        assert err < 1e-5, f"sharded gradient sum differs from the full-minibatch gradient: {err:.2e}"


@pytest.mark.gpu
@pytest.mark.parametrize("precision,exchange,shape", [("fp32", "nccl", ""), ("bf16", "nccl", ""), ("bf16", "p2p", ""), ("bf16", "p2p", "chain"),
                                                      ("bf16", "nccl", "chain"), ("bf16", "p2p", "chain+replicate")])
def test_nccl_data_parallel_matches_single_gpu(precision, exchange, shape):
    """exchange: the per-minibatch gradient sum through ncclAllReduce, or through peer-mapped memory fused into the
    optimizer kernel (bf16 path).  +replicate: the peers' observation tables copied once per rollout
    (b200ppo_table_replicate) instead of gathered remotely every epoch."""
    replicate = shape.endswith("+replicate")
    shape = shape.replace("+replicate", "")
    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29577", os.path.join(ROOT, "tests", "run_dist_parity.py"), precision] + ([shape] if shape else [])
    env = dict(os.environ, B200PPO_P2P="1" if exchange == "p2p" else "0", B200PPO_TABLE_REPLICATE="1" if replicate else "0")
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=300, env=env)
    assert res.returncode == 0 and "DIST PARITY OK" in res.stdout, res.stdout[-2000:] + res.stderr[-2000:]
    assert f"exchange {exchange}" in res.stdout, res.stdout[-2000:]
