"""Seeded synthetic overlap-segment mixtures.

Speech-shaped sources: white noise -> band-limiting one-pole filters (~300-3400 Hz) x slow
amplitude envelope x random on/off gating; each source scaled to an RMS drawn from the
LibriMix loudness range (-33..-25 dBFS), summed and peak-limited to 0.9 exactly like the
reference's mixing tool does (``/root/reference/scripts/tools/mix_wavs.py:139-173``: dry-run
sum in float64, one global attenuation if the peak exceeds the limit, cast to float32).
The sources are the PIT references.  Used by tests and bench.py for BOTH arms, so the CUDA
path and the CPU oracle see identical arrays.
"""
from __future__ import annotations

from typing import List, Tuple

import numpy as np

__all__ = ["make_mixture", "segment_lengths", "make_batch", "streaming_chunk_samples"]


def _one_pole_band(x: np.ndarray, sr: int, f_lo: float = 300.0, f_hi: float = 3400.0) -> np.ndarray:
    from scipy.signal import lfilter
    a_lp = float(np.exp(-2.0 * np.pi * min(f_hi, 0.45 * sr) / sr))
    y = lfilter([1.0 - a_lp], [1.0, -a_lp], x)
    a_hp = float(np.exp(-2.0 * np.pi * f_lo / sr))
    return lfilter([1.0, -1.0], [1.0, -a_hp], y)


def make_mixture(seed: int, n_samples: int, n_src: int = 3, sr: int = 8000,
                 peak_limit: float = 0.9) -> Tuple[np.ndarray, List[np.ndarray]]:
    """Returns ``(mixture fp32 (T,), [n_src source fp32 (T,)])`` for ``seed``."""
    rng = np.random.default_rng(1000 + int(seed))
    t = np.arange(n_samples, dtype=np.float64) / sr
    srcs = []
    for _ in range(n_src):
        x = _one_pole_band(rng.standard_normal(n_samples), sr)
        f_m = rng.uniform(2.0, 6.0)
        env = 0.5 * (1.0 + np.sin(2.0 * np.pi * f_m * t + rng.uniform(0, 2 * np.pi)))
        gate = np.ones(n_samples)
        n_gaps = int(rng.integers(0, 3))
        for _g in range(n_gaps):  # random "silent" spans so the amount of overlap varies
            g_len = int(rng.uniform(0.05, 0.25) * n_samples)
            g0 = int(rng.integers(0, max(1, n_samples - g_len)))
            gate[g0:g0 + g_len] = 0.05
        x = x * env * gate
        rms = float(np.sqrt(np.mean(x ** 2)) + 1e-12)
        x = x * (10.0 ** (rng.uniform(-33.0, -25.0) / 20.0) / rms)
        srcs.append(x)
    mix = np.sum(srcs, axis=0)
    peak = float(np.max(np.abs(mix)) + 1e-12)
    g = peak_limit / peak if peak > peak_limit else 1.0
    return (mix * g).astype(np.float32), [(s * g).astype(np.float32) for s in srcs]


def segment_lengths(n_seg: int, seed: int, lo_s: float = 0.2, hi_s: float = 4.0,
                    sr: int = 8000) -> List[int]:
    """``n_seg`` durations ~ U[lo_s, hi_s] seconds (BASELINE.json configs 2/4/5), in samples."""
    rng = np.random.default_rng(int(seed))
    return [int(round(d * sr)) for d in rng.uniform(lo_s, hi_s, size=n_seg)]


def make_batch(n_seg: int, seed: int, n_src: int = 3, sr: int = 8000, lo_s: float = 0.2,
               hi_s: float = 4.0, with_sources: bool = False):
    """A batch of overlap segments.  Returns ``[mix_i]`` or ``([mix_i], [[src_ij]])``."""
    lens = segment_lengths(n_seg, seed, lo_s, hi_s, sr)
    mixes, sources = [], []
    for i, n in enumerate(lens):
        m, s = make_mixture(seed * 100003 + i, n, n_src, sr)
        mixes.append(m)
        sources.append(s)
    return (mixes, sources) if with_sources else mixes


def streaming_chunk_samples(process_seconds: float, sr: int = 8000, chunk_size: int = 1024) -> int:
    """Chunk length of the reference's streaming path:
    ``frames_per_process = int(sr * process_seconds / chunk_size)`` reads of ``chunk_size``
    samples (``/root/reference/scripts/osd/streaming_overlap_3src.py:89-91``)."""
    return int(sr * process_seconds / chunk_size) * chunk_size
