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
    """All ranks contribute ``records`` [k_r, S] float64 and ``keys`` [(frame, molecule), ...];
    every rank receives the concatenation in rank order."""
    import torch
    import torch.distributed as dist
    rank, size = world()
    if size == 1:
        return records, keys
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
    pay = torch.zeros((kmax, stride + 2), dtype=torch.float64, device=dev)
    if len(keys):
        rows = np.zeros((len(keys), stride + 2))
        rows[:, 0] = [k[0] for k in keys]
        rows[:, 1] = [-1 if isinstance(k[1], str) else k[1] for k in keys]
        rows[:, 2:] = records
        pay[:len(keys)] = torch.from_numpy(rows).to(dev)
    out = torch.zeros((size * kmax, stride + 2), dtype=torch.float64, device=dev)
    dist.all_gather_into_tensor(out, pay)
    out = out.cpu().numpy().reshape(size, kmax, stride + 2)
    recs, allkeys = [], []
    for r in range(size):
        blk = out[r, :counts[r]]
        recs.append(blk[:, 2:])
        allkeys += [(int(a), "0" if b < 0 else int(b)) for a, b in blk[:, :2]]
    return (np.concatenate(recs) if recs else np.zeros((0, stride))), allkeys
