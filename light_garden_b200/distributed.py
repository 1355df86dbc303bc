"""One process per GPU: the host-side plumbing of the ray-shard + image-reduce scheme.

torch.distributed is used for exactly two things: carrying the 128-byte NCCL unique id from rank 0 to the other
ranks, and barriers/timing in bench.py.  The data path (trace, accumulate, ncclReduce of the partial images) is
inside the C library.
"""
import ctypes as C


def shard_count(n: int, rank: int, world: int) -> int:
    """Number of rays of a light with n rays that rank `rank` of `world` traces: the rays rank, rank + world, ...
    (same interleaved split as lg_shard_set)."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError("rank/world")
    return (n - rank + world - 1) // world if n > rank else 0


def shard_indices(n: int, rank: int, world: int):
    return range(rank, n, world) if shard_count(n, rank, world) else range(0)


def broadcast_bytes(payload: bytes, n: int, rank: int, src: int = 0, device=None) -> bytes:
    """Broadcast `n` bytes from `src` over the default process group (any backend)."""
    import torch
    import torch.distributed as dist
    buf = torch.zeros(n, dtype=torch.uint8)
    if rank == src:
        buf = torch.tensor(list(payload[:n]), dtype=torch.uint8)
    if device is not None:
        buf = buf.to(device)
    dist.broadcast(buf, src)
    return bytes(buf.cpu().tolist())


def init_comm(ctx, rank: int, world: int, device=None):
    """Create the context's NCCL communicator: rank 0 draws the unique id, everybody joins."""
    from ._lib import check, load
    lib = load()
    raw = (C.c_ubyte * 128)()
    if rank == 0:
        check(None, lib.lg_comm_unique_id(raw))
    ident = broadcast_bytes(bytes(raw), 128, rank, 0, device)
    raw = (C.c_ubyte * 128)(*ident)
    ctx.call("lg_comm_init_rank", raw, rank, world)
