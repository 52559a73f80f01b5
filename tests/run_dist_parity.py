"""Multi-GPU parity (launch with torchrun, one rank per GPU):  the data-parallel update over G ranks must equal the
single-GPU update with the same GLOBAL minibatch (fp32 reassociation only).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 tests/run_dist_parity.py
"""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mujoco_reinforcement_learning_b200 as pkg  # noqa: E402
from mujoco_reinforcement_learning_b200 import distributed as D  # noqa: E402


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    precision = sys.argv[1] if len(sys.argv) > 1 else "fp32"
    n_local, T, Dm, A, H, GB = 64, 32, 40, 6, [64, 64], 512
    if len(sys.argv) > 2 and sys.argv[2] == "chain":  # 256-wide nets: every rank's slice of a minibatch goes through tc_chain_kernel
        n_local, T, Dm, A, H, GB = 128, 32, 120, 6, [256, 256], 2048
    N = n_local * world
    M = N * T

    def make_agent():
        run = pkg.Run(training_config=pkg.TrainingConfig(batch_size=GB, epochs_per_iteration=2),
                      environment_config=pkg.EnvironmentConfig(maximum_timesteps=T, num_envs=N),
                      network_config=pkg.NetworkConfig(input_shape=Dm, output_shape=A, linear_hidden_shapes=H, critic_hidden_shapes=H),
                      device=str(dev), gemm_precision=precision)
        torch.manual_seed(0)
        return run, pkg.PPOAgent(run, max_batch=GB)

    g = torch.Generator().manual_seed(7)
    full = {"current_state": torch.randn(N, T, Dm, generator=g), "action": torch.randn(N, T, A, generator=g),
            "action_log_prob": torch.randn(N, T, generator=g) * 0.1 - 8.0, "advantage": torch.randn(N, T, 1, generator=g),
            "current_state_value_target": torch.randn(N, T, 1, generator=g)}
    perms = torch.stack([torch.randperm(M, generator=g) for _ in range(2)])
    # distributed: every rank holds its env slab, all-gathers, consumes its slice of each global minibatch
    run, agent = make_agent()
    D.init_engine_comm(agent.engine)
    sl = slice(rank * n_local, (rank + 1) * n_local)
    mine = {k: v[sl].reshape(n_local * T, -1).to(dev) for k, v in full.items()}
    gathered = D.all_gather_fields(mine)
    for k, v in full.items():
        assert torch.equal(gathered[k].cpu(), v.reshape(M, -1)), f"all-gather order broken for {k}"
    shared = D.share_rollout(agent.engine, mine)  # peer-memory path: observations stay with their owners (bf16 tables)
    if agent.engine.p2p:
        assert shared["current_state"] is None
        gathered = shared
    hp = agent.engine.hparams(1e-4, 1e-4, 0.1, 1e-4)
    init_flat = agent.engine.flat.clone()
    losses_d = agent.engine.train(gathered["current_state"], gathered["action"], gathered["action_log_prob"].reshape(M),
                                  gathered["advantage"].reshape(M), gathered["current_state_value_target"].reshape(M),
                                  perms.to(dev), GB, hp)
    torch.cuda.synchronize()
    # the same update again from the same initial state, every rank uploading only the permutation slots it consumes
    p_first, l_first = agent.engine.flat.clone(), losses_d.clone()
    agent.engine.flat.copy_(init_flat)
    agent.engine.exp_avg.zero_()
    agent.engine.exp_avg_sq.zero_()
    agent.engine.adam_step = 0
    my_perms = D.slice_perms_for_rank(perms.pin_memory(), GB, world, rank, dev)
    assert my_perms.shape == (2, M // GB, GB // world)
    losses_s = agent.engine.train(gathered["current_state"], gathered["action"], gathered["action_log_prob"].reshape(M),
                                  gathered["advantage"].reshape(M), gathered["current_state_value_target"].reshape(M),
                                  my_perms, GB, hp, rank_sliced_perms=True)
    torch.cuda.synchronize()
    sliced_same = torch.equal(agent.engine.flat, p_first) and torch.equal(losses_s, l_first)
    # single GPU reference on this rank: same global minibatch, world = 1
    run1, agent1 = make_agent()
    f = {k: v.reshape(M, -1).to(dev) for k, v in full.items()}
    losses_1 = agent1.engine.train(f["current_state"], f["action"], f["action_log_prob"].reshape(M), f["advantage"].reshape(M),
                                   f["current_state_value_target"].reshape(M), perms.to(dev), GB, hp)
    torch.cuda.synchronize()
    tol = 1e-5 if precision == "fp32" else 2e-2
    p_d, p_1 = agent.engine.flat, agent1.engine.flat
    err_p = ((p_d - p_1).abs().max() / p_1.abs().max()).item()
    err_l = ((losses_d - losses_1).abs().max() / losses_1.abs().max()).item()
    # every rank must hold identical parameters after the all-reduced steps
    ref = p_d.clone()
    dist.broadcast(ref, src=0)
    same = torch.equal(ref, p_d)
    print(f"rank {rank}/{world} [{precision}] exchange {'p2p' if agent.engine.p2p else 'nccl'}: param err {err_p:.2e} loss err {err_l:.2e} replicas identical {same} rank-sliced perms identical {sliced_same}", flush=True)
    ok = err_p <= tol and err_l <= max(tol, 1e-5) and same and sliced_same and agent.engine.adam_step == agent1.engine.adam_step
    flag = torch.tensor([0 if ok else 1], device=dev)
    dist.all_reduce(flag)
    dist.barrier()
    dist.destroy_process_group()
    if flag.item() != 0:
        sys.exit(1)
    if rank == 0:
        print("DIST PARITY OK")


if __name__ == "__main__":
    main()
