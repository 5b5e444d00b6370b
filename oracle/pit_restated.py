"""ORACLE (test infrastructure, never shipped or measured as the product).

NumPy restatement of the reference's PIT SI-SDR / SI-SDRi evaluation.  PARITY PINNED:
``oracle/make_golden.py`` imports the reference's own functions from ``/root/reference``
(in the build container) and stores their outputs on seeded inputs in
``tests/golden/pit_golden.npz``; ``tests/test_oracle_pit.py`` checks this restatement
against those vectors bit-for-bit (same NumPy ops in the same order).

Two variants exist in the reference and both are restated:

* "o3"  -- ``/root/reference/scripts/osd/overlap3_core.py``
    ``_si_sdr`` :48-63, ``_pit_best_si_sdr_k`` :66-100, ``_compute_sdr_improvement_pit_k`` :103-117
* "ews" -- ``/root/reference/scripts/osd/evaluate_with_sources.py``
    ``si_sdr`` :286-305, ``perm_si_sdr`` :308-317, ``pit_best_si_sdr`` :334-374,
    ``compute_sdr_improvement_pit`` :377-393, ``pit_best_si_sdr_k`` :396-428,
    ``compute_sdr_improvement_pit_k`` :431-441

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline legs may import this.
"""
from __future__ import annotations

import itertools
import math
from typing import List, Tuple

import numpy as np


# --------------------------------------------------------------------------- "o3" variant
def si_sdr_o3(reference: np.ndarray, estimation: np.ndarray) -> float:
    """overlap3_core.py:48-63 -- fp32 arrays, fp64 Python scalars, +1e-12 guards."""
    if reference.shape != estimation.shape:
        n = min(reference.shape[-1], estimation.shape[-1])
        reference = reference[..., :n]
        estimation = estimation[..., :n]
    ref = reference.astype(np.float32) - float(np.mean(reference))
    est = estimation.astype(np.float32) - float(np.mean(estimation))
    ref_energy = float(np.sum(ref ** 2)) + 1e-12
    if ref_energy <= 0:
        return float("nan")
    scale = float(np.dot(est, ref)) / ref_energy
    proj = scale * ref
    e_noise = est - proj
    num = float(np.sum(proj ** 2)) + 1e-12
    den = float(np.sum(e_noise ** 2)) + 1e-12
    return 10.0 * float(np.log10(num / den))


def pit_select_o3(sdr_mat: List[List[float]], K: int, N: int) -> Tuple[float, List[int]]:
    """The enumeration of overlap3_core.py:78-100 over a precomputed K x N pair matrix
    (``sdr_mat[i][j] = si_sdr(ref_i, pred_j)``).  Lexicographic combinations x permutations,
    left-to-right sum from 0.0, ``/ float(K)``, strict ``>`` from ``best = -1e9``;
    ``assigned[i] = cols[perm.index(i)]`` = pred index assigned to ref i."""
    best = -1e9
    best_idx: List[int] = []
    for cols in itertools.combinations(range(N), K):
        for perm in itertools.permutations(range(K), K):
            s = 0.0
            valid = True
            for r_i, c_i in enumerate(cols):
                sdr_val = sdr_mat[perm[r_i]][c_i]
                if np.isnan(sdr_val):
                    valid = False
                    break
                s += sdr_val
            if not valid:
                continue
            mean_sdr = s / float(K)
            if mean_sdr > best:
                best = mean_sdr
                best_idx = [cols[perm.index(i)] for i in range(K)]
    if best_idx == []:
        return float("nan"), []
    return float(best), best_idx


def pit_best_si_sdr_k_o3(refs, preds) -> Tuple[float, List[int]]:
    """overlap3_core.py:66-100."""
    K = len(refs)
    if K not in (2, 3):
        raise ValueError("_pit_best_si_sdr_k supports K=2 or 3")
    if len(preds) < K:
        return float("nan"), []
    sdr_mat = [[si_sdr_o3(refs[i], p) for p in preds] for i in range(K)]
    return pit_select_o3(sdr_mat, K, len(preds))


def compute_sdr_improvement_pit_k_o3(mix_chunk, refs, preds) -> Tuple[float, float, List[int]]:
    """overlap3_core.py:103-117."""
    if len(refs) not in (2, 3):
        return float("nan"), float("nan"), []
    base_vals = [si_sdr_o3(r, mix_chunk) for r in refs]
    if any(np.isnan(x) for x in base_vals):
        return float("nan"), float("nan"), []
    base = float(np.mean(base_vals))
    best, indices = pit_best_si_sdr_k_o3(refs, preds)
    if np.isnan(best):
        return float("nan"), float("nan"), []
    return float(best), float(best - base), indices


# --------------------------------------------------------------------------- "ews" variant
def si_sdr_ews(reference: np.ndarray, estimation: np.ndarray) -> float:
    """evaluate_with_sources.py:286-305 -- every scalar stays in the array dtype (np.float32
    for fp32 inputs); the log is ``math.log10`` so the return value is a Python float."""
    if reference.shape != estimation.shape:
        m = min(len(reference), len(estimation))
        reference = reference[:m]
        estimation = estimation[:m]
    ref = reference - np.mean(reference)
    est = estimation - np.mean(estimation)
    ref_energy = np.sum(ref ** 2) + 1e-12
    scale = np.dot(est, ref) / ref_energy
    proj = scale * ref
    e_noise = est - proj
    ratio = (np.sum(proj ** 2) + 1e-12) / (np.sum(e_noise ** 2) + 1e-12)
    return 10 * math.log10(ratio)


def pit2_select_ews(sdr1: List[float], sdr2: List[float]) -> Tuple[float, List[int], bool]:
    """The K=2 search of evaluate_with_sources.py:352-374 over precomputed rows:
    ordered pairs j != k, unswapped preferred on ``>=``, first strictly-greater kept; the
    returned pair is ``[j, k]`` (NOT reordered when swapped) plus the ``swapped`` flag."""
    best = -1e9
    best_pair = (-1, -1)
    best_swapped = False
    P = len(sdr1)
    for j in range(P):
        for k in range(P):
            if j == k:
                continue
            sdr_12 = 0.5 * (sdr1[j] + sdr2[k])
            sdr_21 = 0.5 * (sdr1[k] + sdr2[j])
            if sdr_12 >= sdr_21:
                cur, swapped = sdr_12, False
            else:
                cur, swapped = sdr_21, True
            if cur > best:
                best, best_pair, best_swapped = cur, (j, k), swapped
    return float(best), [best_pair[0], best_pair[1]], best_swapped


def pit_best_si_sdr_ews(refs, preds) -> Tuple[float, List[int], bool]:
    """evaluate_with_sources.py:334-374."""
    assert len(refs) == 2, "This PIT helper expects exactly 2 references."
    if len(preds) < 2:
        return float("nan"), [], False
    s1_ref, s2_ref = refs
    sdr1 = [si_sdr_ews(s1_ref, p) for p in preds]
    sdr2 = [si_sdr_ews(s2_ref, p) for p in preds]
    return pit2_select_ews(sdr1, sdr2)


def compute_sdr_improvement_pit_ews(mix_chunk, s1_ref, s2_ref, preds):
    """evaluate_with_sources.py:377-393."""
    base = 0.5 * (si_sdr_ews(s1_ref, mix_chunk) + si_sdr_ews(s2_ref, mix_chunk))
    best, indices, swapped = pit_best_si_sdr_ews([s1_ref, s2_ref], preds)
    if math.isnan(best):
        return float("nan"), float("nan"), [], False
    return best, best - base, indices, swapped


def pit3_select_ews(sdr_mat: List[List[float]], N: int) -> Tuple[float, List[int]]:
    """The K=3 search of evaluate_with_sources.py:414-428: ``permutations(cols, 3)`` with
    ``perm[i]`` = pred index for ref i, ``(a + b + c) / 3.0``, strict ``>``."""
    best = -1e9
    best_idx: List[int] = []
    for cols in itertools.combinations(range(N), 3):
        for perm in itertools.permutations(cols, 3):
            cur = (sdr_mat[0][perm[0]] + sdr_mat[1][perm[1]] + sdr_mat[2][perm[2]]) / 3.0
            if cur > best:
                best = cur
                best_idx = list(perm)
    return float(best), best_idx


def pit_best_si_sdr_k_ews(refs, preds) -> Tuple[float, List[int]]:
    """evaluate_with_sources.py:396-428."""
    K = len(refs)
    if K == 2:
        best, pair, _ = pit_best_si_sdr_ews(refs, preds)
        return best, pair
    if K != 3:
        raise NotImplementedError("Only K in {2,3} supported")
    if len(preds) < 3:
        return float("nan"), []
    sdr_mat = [[si_sdr_ews(refs[i], p) for p in preds] for i in range(3)]
    return pit3_select_ews(sdr_mat, len(preds))


def compute_sdr_improvement_pit_k_ews(mix_chunk, refs, preds):
    """evaluate_with_sources.py:431-441."""
    if len(refs) not in (2, 3):
        raise NotImplementedError("Only K=2 or 3 refs are supported")
    base = float(np.mean([si_sdr_ews(r, mix_chunk) for r in refs]))
    best, idx = pit_best_si_sdr_k_ews(refs, preds)
    if math.isnan(best):
        return float("nan"), float("nan"), []
    return best, best - base, idx


# --------------------------------------------------------------------------- helpers for tests
def si_sdr_f64(reference: np.ndarray, estimation: np.ndarray) -> float:
    """Float64 SI-SDR (same formula) used to MEASURE parity between two waveforms in tests;
    not part of the reference."""
    r = np.asarray(reference, np.float64)
    e = np.asarray(estimation, np.float64)
    r = r - r.mean()
    e = e - e.mean()
    a = float(np.dot(e, r)) / (float(np.dot(r, r)) + 1e-30)
    p = a * r
    return 10.0 * math.log10((float(np.dot(p, p)) + 1e-30) / (float(np.dot(e - p, e - p)) + 1e-30))


def aggregate_stats(vals: List[float]):
    """overlap3_core.py:860-869 -- float32 array, mean / median / population std."""
    if not vals:
        return None, None, None
    arr = np.asarray(vals, dtype=np.float32)
    return float(np.mean(arr)), float(np.median(arr)), float(np.std(arr))
