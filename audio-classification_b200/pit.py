"""Batched PIT SI-SDR / SI-SDRi on the GPU, behind the reference's own function signatures.

Replaces (``/root/reference``):
  * ``scripts/osd/overlap3_core.py`` ``_si_sdr`` :48-63, ``_pit_best_si_sdr_k`` :66-100,
    ``_compute_sdr_improvement_pit_k`` :103-117        (variant ``"overlap3"``)
  * ``scripts/osd/evaluate_with_sources.py`` ``compute_sdr_improvement_pit`` :377-393,
    ``compute_sdr_improvement_pit_k`` :431-441          (variant ``"evalsrc"``)

``pit_sisdr_batch`` scores a whole list of segments in two kernel launches
(``b200sep_pit_sisdr_batch``); the single-segment wrappers keep the reference call shape so the
evaluation call site (``overlap3_core.py:726-728``) can switch with a one-line import change.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import List, Sequence, Tuple

import numpy as np

from . import _native as nat


@dataclass
class PitResult:
    sisdr: np.ndarray    # [n_seg] float64 best mean SI-SDR (NaN when undefined)
    sisdri: np.ndarray   # [n_seg] float64 best - mean_i SI-SDR(ref_i, mix)
    perm: np.ndarray     # [n_seg, K] int32: estimate index assigned to reference i (-1 when undefined)
    swapped: np.ndarray  # [n_seg] bool, K=2 "evalsrc" variant only
    pair: np.ndarray     # [n_seg, 3, W] float64, W = max(N, 3) + 1: pair[i][j] = SI-SDR(ref_i, est_j); [:, :, -1] = vs mixture
    gap: np.ndarray = None  # [n_seg] float64: best - second-best mean SI-SDR over the distinct assignments (dB)


def pit_sisdr_device(d_refs, d_ests, d_mix, lens: Sequence[int], K: int, N: int,
                     variant: str = "overlap3"):
    """Device-resident entry: all tensors 1-D float32 CUDA, ragged layout of include/b200sep.h
    (segment i, source k at ``K*off_i + k*len_i``).  Returns CUDA tensors
    ``(pair [n,3,W] f64, sisdr [n] f64, sisdri [n] f64, perm [n,4] i32, gap [n] f64)``, W = max(N, 3) + 1 (N <= 8;
    the last column is the SI-SDR against the mixture); asynchronous."""
    import torch
    if variant not in nat.PIT_VARIANTS:
        raise ValueError(f"unknown PIT variant {variant!r}")
    n = len(lens)
    dev = d_mix.device
    for t in (d_refs, d_ests, d_mix):
        if t.dtype != torch.float32 or not t.is_cuda or not t.is_contiguous():
            raise RuntimeError("PIT inputs must be contiguous float32 CUDA tensors")
    total = int(sum(int(v) for v in lens))
    if d_refs.numel() < K * total or d_ests.numel() < N * total or d_mix.numel() < total:
        raise RuntimeError("PIT input tensors are shorter than the ragged layout requires")
    arr = nat.i32_array(lens)
    need = int(nat.lib().b200sep_pit_workspace_bytes_n(arr, n, int(N)))
    if need == 0:
        raise RuntimeError(f"b200sep_pit_workspace_bytes_n: {nat.last_error()}")
    ws = torch.empty(need + 256, dtype=torch.uint8, device=dev)
    ws_ptr = (ws.data_ptr() + 255) // 256 * 256
    pair = torch.empty((n, 3, max(int(N), 3) + 1), dtype=torch.float64, device=dev)
    sisdr = torch.empty(n, dtype=torch.float64, device=dev)
    sisdri = torch.empty(n, dtype=torch.float64, device=dev)
    perm = torch.empty((n, 4), dtype=torch.int32, device=dev)
    gap = torch.empty(n, dtype=torch.float64, device=dev)
    stream = torch.cuda.current_stream(dev).cuda_stream
    with torch.cuda.device(dev):
        nat.check(nat.lib().b200sep_pit_sisdr_batch(
            C.c_void_p(d_refs.data_ptr()), C.c_void_p(d_ests.data_ptr()), C.c_void_p(d_mix.data_ptr()),
            None, arr, n, int(K), int(N), nat.PIT_VARIANTS[variant], C.c_void_p(pair.data_ptr()),
            C.c_void_p(sisdr.data_ptr()), C.c_void_p(sisdri.data_ptr()), C.c_void_p(perm.data_ptr()),
            C.c_void_p(gap.data_ptr()), C.c_void_p(ws_ptr), C.c_size_t(need), C.c_void_p(stream)), "b200sep_pit_sisdr_batch")
    ws.record_stream(torch.cuda.current_stream(dev))
    return pair, sisdr, sisdri, perm, gap


def _pack(list_of_lists: Sequence[Sequence[np.ndarray]], lens: Sequence[int], width: int) -> np.ndarray:
    """[[src arrays] per segment] -> ragged 1-D float32 (segment-major, then source, then time)."""
    out = np.empty(int(sum(lens)) * width, dtype=np.float32)
    o = 0
    for arrs, n in zip(list_of_lists, lens):
        for a in arrs:
            out[o:o + n] = np.asarray(a, dtype=np.float32).reshape(-1)[:n]
            o += n
    return out


def _select_host(M: np.ndarray, base_vals: np.ndarray, K: int, N: int, variant: str):
    """Host replay of the device finalize step (same enumeration / tie-break order as the reference:
    ``overlap3_core.py:78-97`` or ``evaluate_with_sources.py:352-373, 418-427``) for the rare segment
    whose arrays have different lengths and therefore needed several device passes."""
    import itertools
    best, idx, swapped = -1e9, [], False
    if variant == "overlap3":
        for cols in itertools.combinations(range(N), K):
            for perm in itertools.permutations(range(K), K):
                vals = [M[perm[r], cols[r]] for r in range(K)]
                if any(np.isnan(v) for v in vals):
                    continue
                s = 0.0
                for v in vals:
                    s += float(v)
                mean = s / float(K)
                if mean > best:
                    best, idx = mean, [cols[perm.index(i)] for i in range(K)]
        base = float(np.mean([float(v) for v in base_vals[:K]]))
    elif K == 2:
        for j in range(N):
            for k in range(N):
                if j == k:
                    continue
                s12, s21 = 0.5 * (M[0, j] + M[1, k]), 0.5 * (M[0, k] + M[1, j])
                cur, sw = (s12, False) if s12 >= s21 else (s21, True)
                if cur > best:
                    best, idx, swapped = float(cur), [j, k], sw
        base = 0.5 * (float(base_vals[0]) + float(base_vals[1]))
    else:
        for cols in itertools.combinations(range(N), 3):
            for perm in itertools.permutations(cols, 3):
                cur = (M[0, perm[0]] + M[1, perm[1]] + M[2, perm[2]]) / 3.0
                if cur > best:
                    best, idx = float(cur), list(perm)
        base = float(np.mean([float(v) for v in base_vals[:3]]))
    if not idx or np.isnan(base):
        return float("nan"), float("nan"), [-1] * K, False
    return best, best - base, idx, swapped


def pit_sisdr_batch(mixes: Sequence[np.ndarray], refs: Sequence[Sequence[np.ndarray]],
                    preds: Sequence[Sequence[np.ndarray]], variant: str = "overlap3",
                    device: int = 0) -> PitResult:
    """Host entry.  ``refs[i]`` has K (2 or 3) arrays, ``preds[i]`` N >= K arrays (same K, N for the
    whole batch).  Every reference ``si_sdr`` truncates ITS two arguments to their common length
    (``overlap3_core.py:49-52``).  At the reference call site all arrays of a segment are cut with
    the same sample indices, so one device pass per segment suffices; a segment whose arrays differ
    in length is scored with one device pass per distinct truncation length and assembled on the
    host, which reproduces the pairwise truncation exactly."""
    import torch
    n = len(mixes)
    if n == 0:
        z = np.zeros(0)
        return PitResult(z, z, np.zeros((0, 0), np.int32), np.zeros(0, bool), np.zeros((0, 3, 4)))
    K, N = len(refs[0]), len(preds[0])
    W = max(N, 3) + 1
    if any(len(r) != K for r in refs) or any(len(p) != N for p in preds):
        raise ValueError("all segments of a batch must have the same K and N")
    # pseudo-segments: (segment index, truncation length)
    jobs, ragged = [], {}
    for i, (m, r, p) in enumerate(zip(mixes, refs, preds)):
        lr = [int(np.asarray(a).shape[-1]) for a in r]
        lp = [int(np.asarray(a).shape[-1]) for a in p]
        lm = int(np.asarray(m).shape[-1])
        if len(set(lr + lp + [lm])) == 1:
            jobs.append((i, lm))
        else:
            need = sorted({min(a, b) for a in lr for b in lp} | {min(a, lm) for a in lr})
            ragged[i] = {"lr": lr, "lp": lp, "lm": lm, "jobs": {}}
            for ln in need:
                ragged[i]["jobs"][ln] = len(jobs)
                jobs.append((i, ln))
    lens = [ln for _, ln in jobs]

    def cut(a, ln):
        a = np.asarray(a, dtype=np.float32).reshape(-1)
        return a[:ln] if a.shape[0] >= ln else np.concatenate([a, np.zeros(ln - a.shape[0], np.float32)])

    h_mix = _pack([[cut(mixes[i], ln)] for i, ln in jobs], lens, 1)
    h_ref = _pack([[cut(a, ln) for a in refs[i]] for i, ln in jobs], lens, K)
    h_est = _pack([[cut(a, ln) for a in preds[i]] for i, ln in jobs], lens, N)
    dev = torch.device("cuda", device)
    with torch.cuda.device(dev):
        d_mix = torch.from_numpy(h_mix).to(dev)
        d_ref = torch.from_numpy(h_ref).to(dev)
        d_est = torch.from_numpy(h_est).to(dev)
        pair, sisdr, sisdri, perm, gap = pit_sisdr_device(d_ref, d_est, d_mix, lens, K, N, variant)
        pair, sisdr, sisdri, perm, gap = (pair.cpu().numpy(), sisdr.cpu().numpy(), sisdri.cpu().numpy(), perm.cpu().numpy(),
                                          gap.cpu().numpy())
    o_pair = np.full((n, 3, W), np.nan)
    o_sisdr, o_sisdri = np.full(n, np.nan), np.full(n, np.nan)
    o_perm, o_sw = np.full((n, K), -1, np.int32), np.zeros(n, bool)
    o_gap = np.full(n, np.nan)
    for j, (i, _ln) in enumerate(jobs):
        if i not in ragged:
            o_pair[i], o_sisdr[i], o_sisdri[i] = pair[j], sisdr[j], sisdri[j]
            o_perm[i], o_sw[i] = perm[j, :K], bool(perm[j, 3])
            o_gap[i] = gap[j]
    for i, info in ragged.items():
        M = np.full((3, W), np.nan)
        for a in range(K):
            for b in range(N):
                M[a, b] = pair[info["jobs"][min(info["lr"][a], info["lp"][b])], a, b]
            M[a, W - 1] = pair[info["jobs"][min(info["lr"][a], info["lm"])], a, W - 1]
        best, sdri, idx, sw = _select_host(M[:, :W - 1], M[:, W - 1], K, N, variant)
        o_pair[i], o_sisdr[i], o_sisdri[i], o_perm[i], o_sw[i] = M, best, sdri, idx, sw
    return PitResult(o_sisdr, o_sisdri, o_perm, o_sw, o_pair, o_gap)


# ---- reference-shaped single-segment wrappers ---------------------------------------------------
def _compute_sdr_improvement_pit_k(mix_chunk: np.ndarray, refs: List[np.ndarray],
                                   preds: List[np.ndarray]) -> Tuple[float, float, List[int]]:
    """Same contract as ``overlap3_core._compute_sdr_improvement_pit_k`` (:103-117)."""
    if len(refs) not in (2, 3) or len(preds) < len(refs):
        return float("nan"), float("nan"), []
    r = pit_sisdr_batch([mix_chunk], [refs], [preds], "overlap3")
    if np.isnan(r.sisdr[0]) or np.isnan(r.sisdri[0]):
        return float("nan"), float("nan"), []
    return float(r.sisdr[0]), float(r.sisdri[0]), [int(v) for v in r.perm[0]]


def _pit_best_si_sdr_k(refs: List[np.ndarray], preds: List[np.ndarray]) -> Tuple[float, List[int]]:
    """Same contract as ``overlap3_core._pit_best_si_sdr_k`` (:66-100)."""
    K = len(refs)
    if K not in (2, 3):
        raise ValueError("_pit_best_si_sdr_k supports K=2 or 3")
    if len(preds) < K:
        return float("nan"), []
    r = pit_sisdr_batch([refs[0]], [refs], [preds], "overlap3")
    if np.isnan(r.sisdr[0]):
        return float("nan"), []
    return float(r.sisdr[0]), [int(v) for v in r.perm[0]]


def _si_sdr(reference: np.ndarray, estimation: np.ndarray) -> float:
    """Same contract as ``overlap3_core._si_sdr`` (:48-63), evaluated on the device."""
    r = pit_sisdr_batch([reference], [[reference, reference]], [[estimation, estimation]], "overlap3")
    return float(r.pair[0, 0, 0])


def compute_sdr_improvement_pit(mix_chunk, s1_ref, s2_ref, preds):
    """Same contract as ``evaluate_with_sources.compute_sdr_improvement_pit`` (:377-393)."""
    if len(preds) < 2:
        return float("nan"), float("nan"), [], False
    r = pit_sisdr_batch([mix_chunk], [[s1_ref, s2_ref]], [preds], "evalsrc")
    return float(r.sisdr[0]), float(r.sisdri[0]), [int(v) for v in r.perm[0]], bool(r.swapped[0])


def compute_sdr_improvement_pit_k(mix_chunk, refs, preds):
    """Same contract as ``evaluate_with_sources.compute_sdr_improvement_pit_k`` (:431-441)."""
    if len(refs) not in (2, 3):
        raise NotImplementedError("Only K=2 or 3 refs are supported")
    if len(preds) < len(refs):
        return float("nan"), float("nan"), []
    r = pit_sisdr_batch([mix_chunk], [refs], [preds], "evalsrc")
    return float(r.sisdr[0]), float(r.sisdri[0]), [int(v) for v in r.perm[0]]
