"""ORACLE (test infrastructure, never imported by the product): NumPy restatements of the code either
side of the separation path.

* ``ensure_sr_linear``      -- ``Separator._ensure_sr``, /root/reference/src/osd/separation.py:91-103
* ``resample_sinc``         -- ``torchaudio.functional.resample`` with its defaults, which the reference's loaders
                               call at /root/reference/scripts/osd/overlap3_core.py:136-143 and
                               /root/reference/scripts/osd/evaluate_with_sources.py:163-178.  torchaudio is a
                               third-party dependency (2.11 in this image); its published algorithm
                               (``_get_sinc_resample_kernel`` / ``_apply_sinc_resample_kernel``) is restated here in
                               fp32 NumPy and PINNED against torchaudio itself (tests/golden/aux_golden.npz, made by
                               oracle/make_golden_aux.py).
* ``frame_rms``             -- /root/reference/scripts/osd/evaluate_with_sources.py:181-197
* ``build_gt_overlap_mask`` -- /root/reference/scripts/osd/evaluate_with_sources.py:226-241
  (both PINNED against the reference's own functions imported through oracle/shim.py).
"""
from __future__ import annotations

import math

import numpy as np


def ensure_sr_linear(samples: np.ndarray, sr: int, sample_rate: int) -> np.ndarray:
    """separation.py:91-103."""
    if sr == sample_rate:
        return samples
    if len(samples) <= 1:
        return samples
    tgt_n = int(round(len(samples) * sample_rate / sr))
    if tgt_n <= 1:
        return samples
    old_idx = np.arange(len(samples), dtype=np.float64)
    new_idx = np.linspace(0, len(samples) - 1, tgt_n, dtype=np.float64)
    return np.interp(new_idx, old_idx, samples).astype(np.float32)


def sinc_kernel_bank(orig_freq: int, new_freq: int, lowpass_filter_width: int = 6, rolloff: float = 0.99):
    """torchaudio ``_get_sinc_resample_kernel`` (sinc_interp_hann) evaluated in float32, as ``resample`` does for a
    float32 waveform.  Returns (bank [new][2 * width + orig] float32, width, orig, new) on the gcd-reduced rates."""
    g = math.gcd(int(orig_freq), int(new_freq))
    orig, new = int(orig_freq) // g, int(new_freq) // g
    base_freq = min(orig, new) * rolloff
    width = math.ceil(lowpass_filter_width * orig / base_freq)
    f32 = np.float32
    idx = np.arange(-width, width + orig, dtype=f32)[None, :] / f32(orig)
    t = np.arange(0, -new, -1, dtype=f32)[:, None] / f32(new) + idx
    t = t * f32(base_freq)
    t = np.clip(t, f32(-lowpass_filter_width), f32(lowpass_filter_width))
    window = np.cos(t * f32(math.pi) / f32(lowpass_filter_width) / f32(2)) ** 2
    t = t * f32(math.pi)
    scale = f32(base_freq / orig)
    with np.errstate(invalid="ignore", divide="ignore"):
        kern = np.where(t == 0, f32(1.0), np.sin(t) / t).astype(f32)
    kern = kern * (window * scale)
    return kern.astype(f32), width, orig, new


def resample_sinc(wav: np.ndarray, orig_freq: int, new_freq: int) -> np.ndarray:
    """``AF.resample(wav[None], orig_freq, new_freq)[0]`` for a 1-D float32 signal (pass-through for equal rates or
    <= 1 sample, as the call sites guard)."""
    wav = np.asarray(wav, dtype=np.float32)
    if orig_freq == new_freq or wav.size <= 1:
        return wav
    bank, width, orig, new = sinc_kernel_bank(orig_freq, new_freq)
    n = wav.shape[0]
    padded = np.concatenate([np.zeros(width, np.float32), wav, np.zeros(width + orig, np.float32)])
    kw = bank.shape[1]
    n_fr = (padded.shape[0] - kw) // orig + 1
    frames = np.lib.stride_tricks.as_strided(padded, shape=(n_fr, kw), strides=(padded.strides[0] * orig, padded.strides[0]))
    out = (frames.astype(np.float64) @ bank.T.astype(np.float64)).astype(np.float32)  # [n_fr][new]
    target = int(math.ceil(new * n / orig))
    return out.reshape(-1)[:target]


def frame_rms(wav: np.ndarray, sr: int, win: float, hop: float) -> np.ndarray:
    """evaluate_with_sources.py:181-197."""
    win_s, hop_s = int(win * sr), int(hop * sr)
    if win_s <= 0:
        raise ValueError("win too small")
    if hop_s <= 0:
        raise ValueError("hop too small")
    out = []
    for start in range(0, max(len(wav) - win_s + 1, 1), hop_s):
        seg = wav[start:start + win_s]
        out.append(0.0 if len(seg) == 0 else float(np.sqrt(np.mean(seg ** 2) + 1e-12)))
    return np.asarray(out, dtype=np.float32)


def build_gt_overlap_mask(s1, s2, sr, win, hop, thr_ratio) -> np.ndarray:
    """evaluate_with_sources.py:226-241."""
    rms1, rms2 = frame_rms(s1, sr, win, hop), frame_rms(s2, sr, win, hop)
    peak = max(rms1.max(initial=0.0), rms2.max(initial=0.0), 1e-9)
    return (rms1 > peak * thr_ratio) & (rms2 > peak * thr_ratio)
