"""Sharding independent 30 s chunks over the GPUs of one box (one process per GPU).

The reference's data parallelism is whisper_full_parallel (/root/reference/thirdparty/whisper.cpp/whisper.cpp:5817-5930):
the audio is cut into pieces, every worker owns a whisper_state, the weights are shared read-only and the results are
concatenated on the host.  Here a worker is a rank with its own B200:

  * the model file is read ONCE (rank 0) and broadcast — the only collective of the whole path (NCCL over NVLink on the
    GPU box, gloo in the CPU tests);
  * chunk i belongs to rank i mod world (round robin keeps the per-rank audio length balanced when clip lengths drift);
  * no collective on the step path; transcripts are gathered on the host at the end.

torch.distributed is plumbing only: nothing numeric happens here.
"""
from __future__ import annotations

import numpy as np


def shard_indices(n_chunks: int, rank: int, world: int) -> list[int]:
    """Chunk ids owned by `rank`: i mod world == rank."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError(f"bad rank/world {rank}/{world}")
    return list(range(rank, n_chunks, world))


def broadcast_model(blob: bytes | None, dist=None, device: str = "cpu", src: int = 0) -> bytes:
    """One broadcast of the ggml model file from `src` to every rank.  `dist` is torch.distributed (initialised) or None."""
    if dist is None or dist.get_world_size() == 1:
        if blob is None:
            raise ValueError("single process: the model bytes must be given")
        return blob
    import torch
    rank = dist.get_rank()
    n = torch.tensor([len(blob) if rank == src else 0], dtype=torch.int64, device=device)
    dist.broadcast(n, src)
    if rank == src:
        buf = torch.frombuffer(bytearray(blob), dtype=torch.uint8).to(device)
    else:
        buf = torch.empty(int(n.item()), dtype=torch.uint8, device=device)
    dist.broadcast(buf, src)
    if rank == src:
        return blob
    return buf.cpu().numpy().tobytes()


def gather_transcripts(local: dict[int, dict], n_chunks: int, dist=None, dst: int = 0) -> list[dict] | None:
    """local: {chunk id: result dict} of this rank.  Returns the list ordered by chunk id on `dst`, None elsewhere."""
    if dist is None or dist.get_world_size() == 1:
        return [local[i] for i in range(n_chunks)]
    world, rank = dist.get_world_size(), dist.get_rank()
    gathered = [None] * world if rank == dst else None
    dist.gather_object(local, gathered, dst=dst)
    if rank != dst:
        return None
    merged: dict[int, dict] = {}
    for part in gathered:
        merged.update(part)
    missing = [i for i in range(n_chunks) if i not in merged]
    if missing:
        raise RuntimeError(f"chunks {missing} were not transcribed by any rank")
    return [merged[i] for i in range(n_chunks)]


def transcribe_sharded(ctx, params, chunks: list[np.ndarray], dist=None, batch: int = 16) -> list[dict] | None:
    """Every rank transcribes its shard of `chunks` (all ranks hold the same list) through whisper_b200_full_batch in
    groups of `batch`; the ordered results come back on rank 0."""
    rank = dist.get_rank() if dist is not None else 0
    world = dist.get_world_size() if dist is not None else 1
    mine = shard_indices(len(chunks), rank, world)
    local: dict[int, dict] = {}
    for g in range(0, len(mine), batch):
        ids = mine[g:g + batch]
        rc = ctx.full_batch(params, [chunks[i] for i in ids])
        if rc != 0:
            raise RuntimeError(f"whisper_b200_full_batch -> {rc} on rank {rank}")
        for j, i in enumerate(ids):
            local[i] = ctx.chunk_result(j)
    return gather_transcripts(local, len(chunks), dist)
