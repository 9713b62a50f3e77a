"""Frame sharding across ranks (one process per GPU) and the single end-of-run gather.

Frames (and molecules inside a frame) are independent work units with no exchange step, so ranks
take contiguous blocks of the frame list and the only collective is one ``all_gather`` of the
fixed-width result records when the analysis is finished (NCCL over NVLink on GPUs; the same code
runs over ``gloo`` on CPU for tests).  The reference's only parallelism is a per-frame
``multiprocessing.Pool`` (trajectory.py:553-586).
"""
from __future__ import annotations

import numpy as np


def world():
    """(rank, world_size) of the initialised ``torch.distributed`` group, else (0, 1)."""
    try:
        import torch.distributed as dist
    except ImportError:
        return 0, 1
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def shard_range(n, rank, size):
    """Contiguous block [lo, hi) of ``n`` work units for ``rank``: ceil(n/size) per rank."""
    per = -(-n // size) if size > 0 else n
    lo = min(n, rank * per)
    return lo, min(n, lo + per)


def gather_records(records, keys):
    """All ranks contribute ``records`` [k_r, S] float64 and their keys (a ``trajectory.RecordKeys`` or a list of
    ``(frame, molecule)`` pairs); every rank receives the concatenation in rank order.  One fixed-width payload row
    per record (frame, molecule key, record) goes through a single ``all_gather_into_tensor``; no per-record Python."""
    import torch
    import torch.distributed as dist
    from .trajectory import RecordKeys
    rank, size = world()
    if size == 1:
        return records, keys
    if not isinstance(keys, RecordKeys):
        keys = RecordKeys.from_pairs(keys)
    backend = dist.get_backend()
    dev = torch.device("cuda", torch.cuda.current_device()) if backend == "nccl" else torch.device("cpu")
    stride = records.shape[1] if records.ndim == 2 else 0
    meta = torch.tensor([len(keys), stride], dtype=torch.int64, device=dev)
    metas = [torch.zeros_like(meta) for _ in range(size)]
    dist.all_gather(metas, meta)
    counts = [int(m[0]) for m in metas]
    stride = max(int(m[1]) for m in metas)
    kmax = max(counts)
    # payload row: frame index, molecule key (numeric; -1 encodes the string key "0"), then the record
    rows = np.zeros((kmax, stride + 2))
    if len(keys):
        rows[:len(keys), 0] = keys.frames
        rows[:len(keys), 1] = keys.mols
        rows[:len(keys), 2:] = records
    pay = torch.from_numpy(rows).to(dev)
    out = torch.empty((size * kmax, stride + 2), dtype=torch.float64, device=dev)
    dist.all_gather_into_tensor(out, pay)
    out = out.cpu().numpy().reshape(size, kmax, stride + 2)
    valid = np.concatenate([out[r, :counts[r]] for r in range(size)]) if sum(counts) else np.zeros((0, stride + 2))
    return (np.ascontiguousarray(valid[:, 2:]),
            RecordKeys(valid[:, 0].astype(np.int64), valid[:, 1].astype(np.int64)))
