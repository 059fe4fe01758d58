"""Chunk partitioning across the GPUs of one box and the host-side gather of per-chunk .alc blobs.

Chunks are fully independent (FrameEncoder carries configuration only, src/pipeline.rs:334-340), so the path
shards with no data-path collective: chunk c -> rank c mod world.  The only cross-rank step is a
host-side gather of the `.alc` blobs into chunk order; torch.distributed is the plumbing for it (gloo on CPU
tests, nccl ranks use the same object collectives through their CPU side).
"""
from __future__ import annotations

from typing import List, Sequence


def chunks_of_rank(n_chunks: int, rank: int, world: int) -> List[int]:
    """Round-robin assignment: chunk c belongs to rank c % world."""
    if not (0 <= rank < world):
        raise ValueError("rank out of range")
    return list(range(rank, n_chunks, world))


def gather_stream(local_blobs: Sequence[bytes], n_chunks: int, rank: int, world: int, dst: int = 0, group=None):
    """Every rank passes the blobs of chunks_of_rank(n_chunks, rank, world), in that order.
    Returns the full ordered list of `.alc` blobs on rank `dst`, None elsewhere."""
    mine = chunks_of_rank(n_chunks, rank, world)
    if len(local_blobs) != len(mine):
        raise ValueError(f"rank {rank} holds {len(local_blobs)} blobs, expected {len(mine)}")
    if world == 1:
        return list(local_blobs)
    import torch.distributed as dist
    gathered = [None] * world if rank == dst else None
    dist.gather_object(list(local_blobs), gathered, dst=dst, group=group)
    if rank != dst:
        return None
    out: List[bytes] = [b""] * n_chunks
    for r in range(world):
        for blob, c in zip(gathered[r], chunks_of_rank(n_chunks, r, world)):
            out[c] = blob
    return out


def concat_stream(blobs: Sequence[bytes]) -> bytes:
    """A multi-chunk stream is the concatenation of self-delimiting .alc blobs (3138-byte header + the three
    compressed_len fields give each blob's length, src/pipeline.rs:137-148)."""
    return b"".join(blobs)


def split_stream(data: bytes) -> List[bytes]:
    out, off = [], 0
    while off < len(data):
        if len(data) - off < 3138 or data[off:off + 4] != b"ALCC":
            raise ValueError("not an .alc stream")
        n = 3138 + sum(int.from_bytes(data[off + 18 + 1040 * c:off + 22 + 1040 * c], "little") for c in range(3))
        out.append(data[off:off + n])
        off += n
    return out
