"""One process per GPU: contiguous particle shards + communicator bootstrap.

Particles never migrate (the x-grid is tiny and replicated), so sharding is a static split of
the particle index range; the only exchange is the all-reduce of the deposited grid inside
vm_field_solve / vm_vproject."""
from __future__ import annotations

import os
from typing import Tuple

from .core import Context


def shard_bounds(n_total: int, rank: int, world: int) -> Tuple[int, int]:
    """[first, last) of rank's contiguous block; sizes differ by at most one, first ranks get the extras."""
    if not (0 <= rank < world):
        raise ValueError("rank out of range")
    base, extra = divmod(int(n_total), int(world))
    first = rank * base + min(rank, extra)
    return first, first + base + (1 if rank < extra else 0)


def init_distributed_context(device: int | None = None, peer_exchange: bool = True) -> Context:
    """Create the Context for this torchrun rank and wire its NCCL communicator.

    The 128-byte NCCL unique id is broadcast from rank 0 over torch.distributed (any host channel
    would do: a Julia host would use MPI.jl).  Works for WORLD_SIZE == 1 without torch.distributed."""
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0")) if device is None else device
    ctx = Context(local)
    if world > 1:
        import torch
        import torch.distributed as dist
        if not dist.is_initialized():
            dist.init_process_group(backend="gloo")     # host-side channel only; the data path is NCCL inside the library
        uid = Context.unique_id() if rank == 0 else bytes(128)
        t = torch.tensor(list(uid), dtype=torch.uint8)
        dist.broadcast(t, src=0)
        ctx.comm_init(rank, world, bytes(t.tolist()))
        if peer_exchange and world <= 8 and os.environ.get("VLASOV_B200_NO_PEER", "0") != "1":
            # fused deposit + exchange + solve over NVLink peer memory: all-gather the CUDA IPC handles
            mine = torch.tensor(list(ctx.peer_handle()), dtype=torch.uint8)
            allh = [torch.zeros(64, dtype=torch.uint8) for _ in range(world)]
            dist.all_gather(allh, mine)
            ctx.peer_connect(b"".join(bytes(h.tolist()) for h in allh))
    return ctx
