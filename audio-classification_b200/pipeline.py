"""Batched overlap-segment driver.

The reference walks the OSD segments of a mixture one by one and calls ``self.sep.separate(chunk, sr)`` -- one
host->device copy, one batch-1 forward, one device->host copy and one synchronisation per overlap segment -- and,
behind ``--eval-separation``, scores each of them with the scalar PIT functions
(``/root/reference/scripts/osd/overlap3_core.py:604-744``; 2-source variant
``/root/reference/scripts/osd/evaluate_with_sources.py:776-826``).  ``OverlapBatchDriver`` keeps the per-segment
rules of that loop (which segments are separated, how chunks and references are sliced, which rows are emitted,
what is counted) but collects ALL overlap chunks -- of one mixture or of many -- into one ``separate_batch`` call and
one ``pit_sisdr_batch`` call.  SV screening and ASR stay with the caller: ``SepRecord.branches`` is exactly the
``branches = list(pred_wavs)`` they consume (``overlap3_core.py:692``).
"""
from __future__ import annotations

import time
from dataclasses import dataclass, field
from typing import Any, Callable, Dict, List, Optional, Sequence, Tuple

import numpy as np

from . import metrics as M

Segment = Tuple[float, float, bool]  # (start s, end s, is_overlap) -- the OSD output the reference iterates


@dataclass
class MixtureJob:
    """One mixture: waveform at ``sr`` (already resampled, as ``mix_np`` is in the reference), its OSD segments and,
    for ``--eval-separation``, the full-length reference sources (``snp`` of ``overlap3_core.py:716-722``)."""
    mix: np.ndarray
    sr: int
    segments: Sequence[Segment]
    mix_path: str = ""
    refs: Optional[Sequence[np.ndarray]] = None


@dataclass
class SepRecord:
    job: int            # index of the MixtureJob
    s: float
    e: float
    chunk: np.ndarray   # mix_np[int(s * sr):int(e * sr)]
    branches: List[np.ndarray] = field(default_factory=list)  # n_src separated streams
    sisdr: Optional[float] = None
    sisdri: Optional[float] = None
    perm: Optional[List[int]] = None


@dataclass
class DriverResult:
    overlap: List[SepRecord]                 # separated segments, in (job, segment) order
    clean: List[Tuple[int, float, float]]    # (job, s, e) of the segments the reference sends straight to SV / ASR
    sep_sisdr_list: List[float]
    sep_sisdri_list: List[float]
    sep_details_rows: List[List[Any]]
    counters: Dict[str, Any]
    time_sep: float

    def sep_metrics(self) -> Dict[str, Any]:
        """The ``sep_*`` keys of metrics.json (``overlap3_core.py:911-925``)."""
        return M.sep_metrics(self.sep_sisdr_list, self.sep_sisdri_list)


class OverlapBatchDriver:
    """``sep`` needs ``separate_batch(chunks, sr) -> List[List[np.ndarray]]`` (b200sep.Separator);
    ``pit`` is ``b200sep.pit_sisdr_batch`` unless another callable with that signature is injected."""

    def __init__(self, sep, min_overlap_dur: float = 0.4, eval_separation: bool = False,
                 pit: Optional[Callable] = None, max_batch_samples: int = 64 * 1024 * 1024, fused_eval: bool = False):
        self.sep = sep
        self.min_overlap_dur = float(min_overlap_dur)
        self.eval_separation = bool(eval_separation)
        self._pit = pit
        self.max_batch_samples = int(max_batch_samples)
        # fused_eval: segments whose references cover them completely are separated AND scored in one device-resident
        # call (``SeparationEngine.separate_and_score``); ``time_sep`` then includes the PIT kernel
        self.fused_eval = bool(fused_eval) and eval_separation and pit is None and hasattr(sep, "engine")

    # ---- selection rules of the reference loop -------------------------------------------------
    def plan(self, jobs: Sequence[MixtureJob]):
        """overlap3_core.py:604-612: skip ``e - s <= 0``; a segment is separated iff it is an overlap AND at least
        ``min_overlap_dur`` long, everything else is a "clean" segment; chunk = ``mix_np[int(s*sr):int(e*sr)]``."""
        overlap: List[SepRecord] = []
        clean: List[Tuple[int, float, float]] = []
        for j, job in enumerate(jobs):
            for s, e, is_olap in job.segments:
                if e - s <= 0:
                    continue
                if (not is_olap) or (e - s) < self.min_overlap_dur:
                    clean.append((j, s, e))
                    continue
                s_i, e_i = int(s * job.sr), int(e * job.sr)
                overlap.append(SepRecord(job=j, s=s, e=e, chunk=np.asarray(job.mix)[s_i:e_i]))
        return overlap, clean

    def _batches(self, recs: Sequence[SepRecord]):
        cur, acc = [], 0
        for r in recs:
            n = int(r.chunk.shape[0])
            if cur and acc + n > self.max_batch_samples:
                yield cur
                cur, acc = [], 0
            cur.append(r)
            acc += n
        if cur:
            yield cur

    def run(self, jobs: Sequence[MixtureJob]) -> DriverResult:
        overlap, clean = self.plan(jobs)
        counters = {"n_seen_overlap_segments": 0, "total_seen_overlap_audio_sec": 0.0, "total_overlap_audio_sec": 0.0,
                    "n_seen_clean_segments": len(clean),
                    "total_seen_clean_audio_sec": float(sum(e - s for _, s, e in clean))}
        # ---- separation: one forward per batch instead of one per segment (overlap3_core.py:689-691) ------------
        time_sep = 0.0
        by_sr: Dict[int, List[SepRecord]] = {}
        for r in overlap:
            by_sr.setdefault(int(jobs[r.job].sr), []).append(r)
        fused_done = set()
        for sr, recs in by_sr.items():
            if self.fused_eval and sr == getattr(self.sep, "sample_rate", sr):
                # references present, K in {2, 3}, every reference slice as long as the chunk -> one fused call per K
                per_k: Dict[int, List[Tuple[SepRecord, List[np.ndarray]]]] = {}
                for r in recs:
                    refs_full = jobs[r.job].refs
                    if not refs_full or r.chunk.shape[0] == 0 or r.chunk.dtype != np.float32:
                        continue
                    K = 3 if len(refs_full) >= 3 else len(refs_full)
                    if K not in (2, 3) or K > int(getattr(self.sep, "n_src", K)):
                        continue
                    s_i, e_i = int(r.s * sr), int(r.e * sr)
                    refs = [np.asarray(x)[s_i:e_i] for x in list(refs_full)[:K]]
                    if all(x.shape[0] == r.chunk.shape[0] for x in refs):
                        per_k.setdefault(K, []).append((r, refs))
                for K, items in per_k.items():
                    for batch in self._batches([r for r, _ in items]):
                        lookup = {id(r): refs for r, refs in items}
                        t0 = time.time()
                        res, wav = self.sep.engine.separate_and_score([r.chunk for r in batch],
                                                                      [lookup[id(r)] for r in batch], "overlap3")
                        time_sep += time.time() - t0
                        for k, r in enumerate(batch):
                            r.branches = list(wav[k])
                            fused_done.add(id(r))
                            best, sdri = float(res.sisdr[k]), float(res.sisdri[k])
                            if not (np.isnan(best) or np.isnan(sdri)):
                                r.sisdr, r.sisdri, r.perm = best, sdri, [int(v) for v in res.perm[k][:K]]
                recs = [r for r in recs if id(r) not in fused_done]
            for batch in self._batches(recs):
                t0 = time.time()
                outs = self.sep.separate_batch([r.chunk for r in batch], sr)
                time_sep += time.time() - t0
                for r, o in zip(batch, outs):
                    r.branches = list(o)
        for r in overlap:  # overlap3_core.py:693-696
            counters["n_seen_overlap_segments"] += 1
            counters["total_seen_overlap_audio_sec"] += r.e - r.s
            counters["total_overlap_audio_sec"] += r.e - r.s
        # ---- --eval-separation (overlap3_core.py:698-744), batched per K --------------------------------------
        sisdr_l: List[float] = []
        sisdri_l: List[float] = []
        rows: List[List[Any]] = []
        if self.eval_separation:
            groups: Dict[Tuple[int, int], List[Tuple[SepRecord, List[np.ndarray]]]] = {}
            for r in overlap:
                refs_full = jobs[r.job].refs
                if not refs_full or id(r) in fused_done:
                    continue
                K = 3 if len(refs_full) >= 3 else len(refs_full)            # :709
                if K not in (2, 3) or len(r.branches) < K:                   # :710-713 (raises -> row skipped)
                    continue
                s_i, e_i = int(r.s * jobs[r.job].sr), int(r.e * jobs[r.job].sr)  # :719-721
                refs = [np.asarray(x)[s_i:e_i] for x in list(refs_full)[:K]]
                if any(x.shape[0] == 0 for x in refs) or r.chunk.shape[0] == 0:
                    continue  # the reference's np.mean of an empty slice yields NaN -> row dropped (:729)
                groups.setdefault((K, len(r.branches)), []).append((r, refs))
            pit = self._pit
            if pit is None:
                from .pit import pit_sisdr_batch as pit
            for (K, N), items in groups.items():
                res = pit([r.chunk for r, _ in items], [refs for _, refs in items],
                          [[np.asarray(b, dtype=np.float32) for b in r.branches] for r, _ in items], "overlap3")
                for k, (r, _) in enumerate(items):
                    best, sdri = float(res.sisdr[k]), float(res.sisdri[k])
                    if np.isnan(best) or np.isnan(sdri):                     # :729
                        continue
                    r.sisdr, r.sisdri, r.perm = best, sdri, [int(v) for v in res.perm[k][:K]]
            for r in overlap:  # rows in the reference's (mixture, segment) order
                if r.sisdr is None:
                    continue
                sisdr_l.append(r.sisdr)
                sisdri_l.append(r.sisdri)
                K = len(r.perm)
                rows.append(M.sep_detail_row(jobs[r.job].mix_path, r.s, r.e, K, r.sisdr, r.sisdri, r.perm))
        return DriverResult(overlap, clean, sisdr_l, sisdri_l, rows, counters, time_sep)
