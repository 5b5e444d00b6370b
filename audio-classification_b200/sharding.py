"""Multi-GPU sharding of independent overlap segments.

A segment's forward and its PIT score depend on nothing else (the reference calls the path at
batch 1: ``/root/reference/src/osd/separation.py:77-83``), so N GPUs = N replicas of the model, each
fed a disjoint shard.  No waveform ever crosses NVLink; the only collective is one ``all_gather`` of
fixed-size per-segment metric records at the end of an evaluation sweep.
"""
from __future__ import annotations

from typing import Callable, List, Sequence, Tuple

import numpy as np

__all__ = ["lpt_partition", "bucket_by_length", "gather_records", "frames_of", "evaluate_sharded"]


def frames_of(n_samples: int, kernel_size: int = 16, stride: int = 8) -> int:
    return (int(n_samples) - kernel_size) // stride + 1


def _lpt_partition_py(costs: Sequence[int], world_size: int) -> List[List[int]]:
    """Pure-Python statement of the partition rule (the test oracle of the native helper)."""
    import heapq
    c = np.asarray(costs, dtype=np.int64).reshape(-1)
    order = np.lexsort((np.arange(c.size), -c))          # descending cost, ties by ascending index
    heap = [(0, k) for k in range(world_size)]            # (load, rank): the least loaded rank, ties by lowest rank
    shards: List[List[int]] = [[] for _ in range(world_size)]
    cl = c.tolist()
    for i in order.tolist():
        load, r = heapq.heappop(heap)
        shards[r].append(i)
        heapq.heappush(heap, (load + cl[i], r))
    return [sorted(s) for s in shards]


def lpt_partition(costs: Sequence[int], world_size: int) -> List[List[int]]:
    """Deterministic longest-processing-time-first partition.  ``costs[i]`` ~ frames of segment i.
    Returns, per rank, the (ascending) list of segment indices it owns.  Runs in the native library
    (``b200sep_lpt_partition``: every rank of a sweep computes it for ALL segments, a fixed cost in front of its shard)."""
    if world_size < 1:
        raise ValueError("world_size must be >= 1")
    from . import _native as nat
    c = np.ascontiguousarray(np.asarray(costs, dtype=np.int64).reshape(-1))
    owner = np.empty(c.size, dtype=np.int32)
    nat.check(nat.lib().b200sep_lpt_partition(c.ctypes.data, int(c.size), int(world_size), owner.ctypes.data),
              "b200sep_lpt_partition")
    idx = np.argsort(owner, kind="stable")               # ascending index inside every rank
    bounds = np.searchsorted(owner[idx], np.arange(world_size + 1))
    return [idx[bounds[r]:bounds[r + 1]].tolist() for r in range(world_size)]


def bucket_by_length(indices: Sequence[int], lens: Sequence[int], max_samples: int) -> List[List[int]]:
    """Length-bucketed batches of a rank's shard: segments sorted by length, cut into batches of at
    most ``max_samples`` samples (bounds the workspace, keeps similar lengths together so the
    per-kernel tile waves are even).  Deterministic."""
    order = sorted(indices, key=lambda i: (int(lens[i]), i))
    out: List[List[int]] = []
    cur: List[int] = []
    tot = 0
    for i in order:
        n = int(lens[i])
        if cur and tot + n > max_samples:
            out.append(cur)
            cur, tot = [], 0
        cur.append(i)
        tot += n
    if cur:
        out.append(cur)
    return out


def gather_records(local_ids: Sequence[int], local_records: np.ndarray, n_total: int,
                   device=None) -> np.ndarray:
    """all_gather of per-segment metric records.  ``local_records`` is ``[len(local_ids), R]``
    float64 (e.g. sisdr, sisdri, perm...).  Returns ``[n_total, R]`` with every segment's record in
    global order on every rank (NaN where no rank reported).  Uses the initialised
    ``torch.distributed`` backend (NCCL over NVLink on GPUs, gloo in CPU tests); with no process
    group it is the identity scatter."""
    import torch
    import torch.distributed as dist
    local_records = np.asarray(local_records, dtype=np.float64).reshape(len(local_ids), -1)
    R = local_records.shape[1] if local_records.size else 0
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        out = np.full((n_total, R), np.nan)
        if len(local_ids):
            out[np.asarray(local_ids, dtype=np.int64)] = local_records
        return out
    world = dist.get_world_size()
    dev = device if device is not None else ("cuda" if dist.get_backend() == "nccl" else "cpu")
    # fixed-size payload per rank: count, R, then (id, record) rows padded to the max shard size
    n_loc = torch.tensor([len(local_ids), R], dtype=torch.int64, device=dev)
    counts = [torch.zeros(2, dtype=torch.int64, device=dev) for _ in range(world)]
    dist.all_gather(counts, n_loc)
    max_n = int(max(int(c[0]) for c in counts))
    R = int(max(int(c[1]) for c in counts))
    payload = torch.full((max_n, R + 1), float("nan"), dtype=torch.float64, device=dev)
    if len(local_ids):
        payload[:len(local_ids), 0] = torch.as_tensor(np.asarray(local_ids, dtype=np.float64), device=dev)
        payload[:len(local_ids), 1:] = torch.as_tensor(local_records, device=dev)
    gathered = [torch.empty_like(payload) for _ in range(world)]
    dist.all_gather(gathered, payload)
    out = np.full((n_total, R), np.nan)
    for r, g in enumerate(gathered):
        k = int(counts[r][0])
        if k:
            g = g[:k].cpu().numpy()
            out[g[:, 0].astype(np.int64)] = g[:, 1:]
    return out


def evaluate_sharded(mixes: Sequence[np.ndarray], refs: Sequence[Sequence[np.ndarray]], sr: int,
                     separate_batch: Callable, pit_batch: Callable, max_samples: int = 4 << 20,
                     kernel_size: int = 16, stride: int = 8, variant: str = "overlap3", device=None,
                     separate_and_score: Callable = None, lens: Sequence[int] = None, n_refs: int = None,
                     score_stream: Callable = None):
    """Separation + PIT SI-SDR / SI-SDRi of a list of independent segments across the ranks of the initialised
    ``torch.distributed`` group (or this process alone): LPT partition on frame counts, length-bucketed batches per
    rank, ``separate_batch(chunks, sr)`` and ``pit_batch(mixes, refs, preds, variant)`` per batch, then ONE
    ``all_gather`` of the per-segment records ``(sisdr, sisdri, perm...)``.  With ``separate_and_score`` (the engine's
    device-resident ``separate_and_score(chunks, refs, variant, return_waveforms=False)``) the estimates never leave the
    GPU between the forward and the PIT kernel; with ``score_stream`` (the engine's ``score_stream(items, variant)``
    generator) consecutive batches are additionally pipelined, two in flight, so host packing and copies overlap the
    kernels.  Every rank returns the full table:
    ``(sisdr[n], sisdri[n], perm[n, K], seconds_spent_separating_on_this_rank)``.  No waveform leaves its rank.

    ``mixes`` / ``refs`` may be sequences over ALL segments, or -- so that a rank only ever materialises its own shard --
    mappings ``{global index: array(s)}`` holding at least this rank's LPT shard, with ``lens`` (samples of every segment
    of the sweep, in global order) and ``n_refs`` given explicitly.

    This is the evaluation loop of ``/root/reference/scripts/osd/overlap3_core.py:689-744`` (separate, score, collect
    ``sep_sisdr_list`` / ``sep_sisdri_list``) over segments that are already cut, for BASELINE.json configs[3] / [4]."""
    import time
    import torch.distributed as dist
    world = dist.get_world_size() if (dist.is_available() and dist.is_initialized()) else 1
    rank = dist.get_rank() if world > 1 else 0
    if lens is None:
        lens = [int(np.asarray(m).shape[-1]) for m in mixes]
    lens = [int(v) for v in lens]
    n = len(lens)
    K = int(n_refs) if n_refs is not None else (len(refs[0]) if n else 0)
    mine = lpt_partition((np.asarray(lens, dtype=np.int64) - kernel_size) // stride + 1, world)[rank] if n else []
    rec = np.full((len(mine), 2 + K), np.nan)
    pos = {seg: k for k, seg in enumerate(mine)}
    t_sep = 0.0
    batches = bucket_by_length(mine, lens, max_samples)
    if score_stream is not None:
        t0 = time.time()
        items = (([mixes[i] for i in batch], [refs[i] for i in batch]) for batch in batches)
        for batch, (res, _) in zip(batches, score_stream(items, variant)):
            rows = [pos[i] for i in batch]
            rec[rows, 0], rec[rows, 1] = res.sisdr, res.sisdri
            rec[rows, 2:] = np.asarray(res.perm, dtype=np.float64)[:, :K]
        t_sep = time.time() - t0
        batches = []
    for batch in batches:
        t0 = time.time()
        if separate_and_score is not None:
            res, _ = separate_and_score([mixes[i] for i in batch], [refs[i] for i in batch], variant, False)
            t_sep += time.time() - t0
        else:
            outs = separate_batch([mixes[i] for i in batch], sr)
            t_sep += time.time() - t0
            res = pit_batch([mixes[i] for i in batch], [refs[i] for i in batch], outs, variant)
        for k, i in enumerate(batch):
            rec[pos[i], 0], rec[pos[i], 1] = res.sisdr[k], res.sisdri[k]
            rec[pos[i], 2:] = np.asarray(res.perm[k][:K], dtype=np.float64)
    table = gather_records(mine, rec, n, device=device)
    return table[:, 0], table[:, 1], table[:, 2:].astype(np.int64), t_sep
