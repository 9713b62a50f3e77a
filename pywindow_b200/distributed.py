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
    ``(frame, molecule)`` pairs); every rank receives the concatenation in rank order.  Two collectives
    (``all_gather_into_tensor`` of the records and of the integer keys, padded to the largest shard); the valid rows
    are selected on the device, so that the host sees ONE device->host copy per array and no per-record Python."""
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
    metas = torch.zeros((size, 2), dtype=torch.int64, device=dev)
    dist.all_gather_into_tensor(metas.view(-1), meta)
    metas = metas.cpu()
    counts = [int(c) for c in metas[:, 0]]
    stride = int(metas[:, 1].max())
    kmax, k = max(counts), len(keys)
    if kmax == 0:
        return np.zeros((0, stride)), RecordKeys(np.zeros(0, np.int64), np.zeros(0, np.int64))
    # own shard, padded to the largest one (molecule key -1 encodes the string key "0")
    rec = torch.zeros((kmax, stride), dtype=torch.float64, device=dev)
    key = torch.zeros((kmax, 2), dtype=torch.int64, device=dev)
    if k:
        rec[:k] = torch.from_numpy(np.ascontiguousarray(records, dtype=np.float64)).to(dev)
        key[:k] = torch.from_numpy(np.stack([np.asarray(keys.frames, dtype=np.int64),
                                             np.asarray(keys.mols, dtype=np.int64)], axis=1)).to(dev)
    all_rec = torch.empty((size, kmax, stride), dtype=torch.float64, device=dev)
    all_key = torch.empty((size, kmax, 2), dtype=torch.int64, device=dev)
    dist.all_gather_into_tensor(all_rec.view(size * kmax, stride), rec)
    dist.all_gather_into_tensor(all_key.view(size * kmax, 2), key)
    if all(c == kmax for c in counts):
        all_rec, all_key = all_rec.view(size * kmax, stride), all_key.view(size * kmax, 2)
    else:
        all_rec = torch.cat([all_rec[r, :counts[r]] for r in range(size)])
        all_key = torch.cat([all_key[r, :counts[r]] for r in range(size)])
    all_key = _device_to_host(all_key.t().contiguous())
    return _device_to_host(all_rec), RecordKeys(all_key[0], all_key[1])


_BOUNCE_DOUBLES = 1 << 20          # 8 MB per pinned bounce buffer


def _device_to_host(t):
    """numpy copy of a contiguous device tensor.  A result-sized pinned block would cost ~0.8 ms/MB to page-lock and a
    pageable ``.cpu()`` runs at a few GB/s through the driver's staging, so large tensors go through two small pinned
    bounce buffers from the engine pool (8 MB each, allocated once per process) and leave them on two helper threads
    -- the copy into fresh pageable memory is page-fault bound."""
    import torch
    if t.device.type != "cuda" or t.numel() * t.element_size() < (4 << 20):
        return t.cpu().numpy()
    from concurrent.futures import ThreadPoolExecutor
    from .engine import default_engine
    eng = default_engine()
    flat = t.view(-1).view(torch.float64) if t.element_size() == 8 else None
    if flat is None:
        return t.cpu().numpy()
    out = np.empty(flat.numel(), dtype=np.float64)
    bounce = [eng.pinned(("bounce", i), (_BOUNCE_DOUBLES,)) for i in range(2)]
    pending = [None, None]

    def drain(i, lo, m, ev):
        ev.synchronize()
        out[lo:lo + m] = bounce[i][:m].numpy()

    with ThreadPoolExecutor(max_workers=2) as pool:
        for c, lo in enumerate(range(0, flat.numel(), _BOUNCE_DOUBLES)):
            i, m = c % 2, min(_BOUNCE_DOUBLES, flat.numel() - lo)
            if pending[i] is not None:
                pending[i].result()                      # the bounce buffer must have been drained
            bounce[i][:m].copy_(flat[lo:lo + m], non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(torch.cuda.current_stream(t.device))
            pending[i] = pool.submit(drain, i, lo, m, ev)
        for fut in pending:
            if fut is not None:
                fut.result()
    return out.view(np.int64 if t.dtype == torch.int64 else np.float64).reshape(tuple(t.shape))
