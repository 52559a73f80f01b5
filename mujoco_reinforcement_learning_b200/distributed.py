"""Multi-GPU plumbing (one process per GPU, `torch.distributed` for rendezvous, NCCL over NVLink for data).

The update is data-parallel over samples: every rank owns a slab of environments for the
rollout and the GAE scan (no communication: the recurrence and the normalisation run along time only); the
slabs are all-gathered once per rollout so that the reference's GLOBAL permutation stays bit-exact; every rank
then consumes rows [r*B/G, (r+1)*B/G) of each global minibatch and the flat fp32 gradient (+ the two loss
partials) is summed with one NCCL all-reduce per minibatch inside `b200ppo_train`.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Dict

import torch
import torch.distributed as dist

from . import _lib


def rank_rows(minibatch_index: int, global_batch: int, world: int, rank: int) -> range:
    """Permutation slots of global minibatch `minibatch_index` consumed by `rank` (same arithmetic as the
    chunked gather in csrc/gather.cu)."""
    lb = global_batch // world
    start = minibatch_index * global_batch + rank * lb
    return range(start, start + lb)


def slice_perms_for_rank(perms: torch.Tensor, global_batch: int, world: int, rank: int, device=None) -> torch.Tensor:
    """The permutation slots rank `rank` consumes, [epochs, M // global_batch, global_batch // world], from the global
    permutations [epochs, M] (`rank_rows` of every minibatch; the dropped tail is left behind).  With `perms` in pinned
    host memory and a CUDA `device` every (epoch, minibatch) piece is one contiguous asynchronous copy on the current
    stream, so a rank uploads 1 / world of the index bytes and the host never touches them
    (`engine.train(..., rank_sliced_perms=True)` consumes the result)."""
    perms = perms.reshape(perms.shape[0], -1) if perms.dim() > 1 else perms.reshape(1, -1)
    epochs, M = perms.shape
    lb = global_batch // world
    nb = M // global_batch
    src = perms[:, :nb * global_batch].reshape(epochs, nb, world, lb)[:, :, rank]
    out = torch.empty((epochs, nb, lb), dtype=perms.dtype, device=device if device is not None else perms.device)
    for e in range(epochs):
        for i in range(nb):
            out[e, i].copy_(src[e, i], non_blocking=True)
    return out


def bind_host_to_gpu(device) -> bool:
    """Pin the calling process to the CPUs NVML reports as local to `device` (its NUMA node), restricted to the CPUs
    the process is already allowed on.  Called once per rank BEFORE the pinned host buffers are allocated: with eight
    ranks on a two-socket host a slab pinned on the far socket crosses the inter-socket link on every host->device copy.
    Returns False (and changes nothing) when NVML is unavailable or fewer than four local CPUs are allowed."""
    try:
        import pynvml
        dev = torch.device(device)
        props = torch.cuda.get_device_properties(dev)
        bus = "%08x:%02x:%02x.0" % (props.pci_domain_id, props.pci_bus_id, props.pci_device_id)
        pynvml.nvmlInit()
        try:
            handle = pynvml.nvmlDeviceGetHandleByPciBusId(bus.encode())
            words = pynvml.nvmlDeviceGetCpuAffinity(handle, ((os.cpu_count() or 1) + 63) // 64)
        finally:
            pynvml.nvmlShutdown()
        local = {64 * i + b for i, w in enumerate(words) for b in range(64) if (int(w) >> b) & 1}
        allowed = os.sched_getaffinity(0)
        both = local & allowed
        if len(both) < 4:  # nothing local is allowed (or too little to hold the launch, NCCL proxy and runtime threads)
            return False
        if both != allowed:
            os.sched_setaffinity(0, both)
        return True
    except Exception:  # affinity is an optimisation: never a reason to fail a run
        return False


def init_engine_comm(engine, group=None) -> None:
    """Create the engine's NCCL communicator: rank 0 makes the unique id, torch.distributed carries it."""
    if not dist.is_initialized():
        raise RuntimeError("torch.distributed is not initialised")
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    lib = _lib.load()
    buf = (C.c_uint8 * 128)()
    if rank == 0:
        _lib.check(lib.b200ppo_comm_unique_id(buf), "b200ppo_comm_unique_id")
    backend = dist.get_backend(group)
    dev = engine.device if backend == "nccl" else torch.device("cpu")
    t = torch.tensor(list(buf), dtype=torch.uint8, device=dev)
    dist.broadcast(t, src=0, group=group)
    ident = (C.c_uint8 * 128)(*t.cpu().tolist())
    with torch.cuda.device(engine.device):
        _lib.check(lib.b200ppo_comm_init(engine._ctx, ident, rank, world), "b200ppo_comm_init")
    engine.rank, engine.world = rank, world
    engine.p2p = False
    # gradient exchange over peer-mapped memory fused into the optimizer kernel (bf16 path, one NVSwitch node)
    if (backend == "nccl" and 2 <= world <= 8 and os.environ.get("B200PPO_P2P", "1") != "0"
            and getattr(engine, "precision", None) == _lib.PREC_BF16):
        # every collective below is executed by every rank whatever happens locally; the exchange is switched on only if
        # all ranks mapped all peers (cudaIpc needs peer access between the devices), otherwise everyone keeps NCCL
        handle = (C.c_uint8 * 64)()
        ok = True
        try:
            with torch.cuda.device(engine.device):
                _lib.check(lib.b200ppo_p2p_export(engine._ctx, handle), "b200ppo_p2p_export")
        except RuntimeError:
            ok = False
        mine = torch.tensor(list(handle), dtype=torch.uint8, device=engine.device)
        everyone = torch.empty(world * 64, dtype=torch.uint8, device=engine.device)
        dist.all_gather_into_tensor(everyone, mine, group=group)
        if ok:
            try:
                handles = (C.c_uint8 * (world * 64))(*everyone.cpu().tolist())
                with torch.cuda.device(engine.device):
                    _lib.check(lib.b200ppo_p2p_import(engine._ctx, handles, world), "b200ppo_p2p_import")
            except RuntimeError:
                ok = False
        agreed = torch.tensor([1 if ok else 0], dtype=torch.int32, device=engine.device)
        dist.all_reduce(agreed, op=dist.ReduceOp.MIN, group=group)  # also: nobody starts an exchange before every rank has mapped its peers
        if int(agreed.item()) == 1:
            with torch.cuda.device(engine.device):
                _lib.check(lib.b200ppo_p2p_enable(engine._ctx, 1), "b200ppo_p2p_enable")
            engine.p2p = True


def all_gather_fields(fields: Dict[str, torch.Tensor], group=None) -> Dict[str, torch.Tensor]:
    """All-gather every leaf along dim 0 (rank-major = env-major: rank r's envs land at rows [r*N_local, ...))."""
    world = dist.get_world_size(group)
    out = {}
    for k, v in fields.items():
        v = v.contiguous()
        full = torch.empty((world * v.shape[0], *v.shape[1:]), dtype=v.dtype, device=v.device)
        dist.all_gather_into_tensor(full, v, group=group)
        out[k] = full
    return out


def share_rollout(engine, fields: Dict[str, torch.Tensor], group=None) -> Dict[str, torch.Tensor]:
    """Make this rank's slab of the rollout available to every rank for the global-permutation update.

    With the peer-memory path (`engine.p2p`): the observations (95 % of the bytes) are NOT all-gathered — every rank
    converts its slab into a bf16 table that its peers have mapped, and `engine.train(None, ...)` pulls the rows of the
    global permutation from the owners over NVLink, on a side stream behind the previous epoch's kernels.  Only the small
    leaves (action, log-prob, advantage, target) are all-gathered.  Returns the fields for `engine.train`, with
    `current_state` = None.  Without it: `all_gather_fields`.
    """
    if not getattr(engine, "p2p", False) or getattr(engine, "_tables_ok", None) is False:
        return all_gather_fields(fields, group)
    world = dist.get_world_size(group)
    lib = _lib.load()
    obs = _lib.require_cuda(fields["current_state"], "current_state", torch.float32).contiguous()
    rows = obs.shape[0]
    obs = obs.reshape(rows, -1)
    # every peer has finished gathering from the tables of the previous rollout once this collective completes
    token = torch.zeros(1, device=obs.device)
    dist.all_reduce(token, group=group)
    if getattr(engine, "_shared_rows", 0) != rows:
        # (re)map the tables; every rank runs every collective, and the tables are used only if all ranks mapped all peers
        handle = (C.c_uint8 * 64)()
        ok = True
        try:
            with torch.cuda.device(engine.device):
                _lib.check(lib.b200ppo_table_export(engine._ctx, rows, handle), "b200ppo_table_export")
        except RuntimeError:
            ok = False
        mine = torch.tensor(list(handle), dtype=torch.uint8, device=engine.device)
        everyone = torch.empty(world * 64, dtype=torch.uint8, device=engine.device)
        dist.all_gather_into_tensor(everyone, mine, group=group)
        if ok:
            try:
                handles = (C.c_uint8 * (world * 64))(*everyone.cpu().tolist())
                with torch.cuda.device(engine.device):
                    _lib.check(lib.b200ppo_table_import(engine._ctx, handles, world, rows), "b200ppo_table_import")
            except RuntimeError:
                ok = False
        agreed = torch.tensor([1 if ok else 0], dtype=torch.int32, device=engine.device)
        dist.all_reduce(agreed, op=dist.ReduceOp.MIN, group=group)
        engine._tables_ok = int(agreed.item()) == 1
        if not engine._tables_ok:
            return all_gather_fields(fields, group)
        engine._shared_rows = rows
    with torch.cuda.device(engine.device):
        _lib.check(lib.b200ppo_table_fill(engine._ctx, _lib.ptr(obs), rows, _lib.stream_ptr()), "b200ppo_table_fill")
    # the all-gather of the small leaves follows the fill in stream order on every rank: when it completes here, every
    # rank's table is complete
    out = all_gather_fields({k: v for k, v in fields.items() if k != "current_state"}, group)
    out["current_state"] = None
    if os.environ.get("B200PPO_TABLE_REPLICATE", "0") == "1":
        # one copy of the peers' tables per rollout instead of (1 - 1/world) of every epoch's rows over NVLink: the epoch
        # gathers become local reads.  Off by default — measured: the remote gathers already hide behind the previous
        # epoch's kernels, while this copy sits on the critical path (2 GPUs 455 -> 460 M samples/s, 8 GPUs 1480 -> 1428 M)
        with torch.cuda.device(engine.device):
            _lib.check(lib.b200ppo_table_replicate(engine._ctx, _lib.stream_ptr()), "b200ppo_table_replicate")
    return out
