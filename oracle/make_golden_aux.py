"""Generates tests/golden/aux_golden.npz (run in the BUILD CONTAINER): outputs of

  * the reference's own ``Separator._ensure_sr`` (/root/reference/src/osd/separation.py:91-103, unmodified, through
    oracle/shim.py),
  * ``torchaudio.functional.resample`` (the resampler the reference's loaders call, overlap3_core.py:136-143),
  * the reference's own ``frame_rms`` / ``build_gt_overlap_mask`` (evaluate_with_sources.py:181-241, unmodified)

on seeded signals.  They pin oracle/aux_restated.py and the CUDA kernels behind b200sep.aux.

    python -m oracle.make_golden_aux
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import shim  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden", "aux_golden.npz")

LINEAR_CASES = [(16000, 8000, 4001), (8000, 16000, 2500), (44100, 8000, 9000), (8000, 8000, 777), (48000, 16000, 1),
                (22050, 16000, 3333), (16000, 8000, 2), (8000, 16000, 3)]
SINC_CASES = [(16000, 8000, 4001), (8000, 16000, 2500), (44100, 16000, 9000), (48000, 8000, 6000), (22050, 16000, 3333)]
RMS_CASES = [(8000, 0.025, 0.010, 16000), (16000, 0.05, 0.02, 9000), (8000, 0.025, 0.010, 150), (8000, 0.025, 0.010, 200),
             (8000, 0.1, 0.05, 12345)]


def signal(seed: int, n: int) -> np.ndarray:
    rng = np.random.default_rng(seed)
    t = np.arange(n) / 8000.0
    env = 0.5 * (1 + np.sin(2 * np.pi * rng.uniform(1, 5) * t + rng.uniform(0, 6)))
    gate = ((np.sin(2 * np.pi * rng.uniform(0.3, 1.5) * t + rng.uniform(0, 6)) > -0.2) | (np.arange(n) < n // 4)).astype(np.float64)
    return (rng.standard_normal(n) * env * gate * 0.1).astype(np.float32)


def main():
    if not shim.reference_available():
        raise SystemExit("needs /root/reference")
    import torchaudio.functional as AF
    RefSep = shim.import_reference_separator()
    _, ews = shim.import_reference_pit()
    out = {}
    sep = RefSep.__new__(RefSep)  # _ensure_sr only reads self.sample_rate; no model is built
    for ci, (sr, tgt, n) in enumerate(LINEAR_CASES):
        x = signal(100 + ci, n)
        sep.sample_rate = tgt
        out[f"lin{ci}_in"] = x
        out[f"lin{ci}_out"] = np.asarray(sep._ensure_sr(x, sr))
        out[f"lin{ci}_cfg"] = np.array([sr, tgt], np.int64)
    for ci, (sr, tgt, n) in enumerate(SINC_CASES):
        x = signal(200 + ci, n)
        y = AF.resample(torch.from_numpy(x).unsqueeze(0), sr, tgt).squeeze(0).numpy()
        out[f"sinc{ci}_in"] = x
        out[f"sinc{ci}_out"] = y
        out[f"sinc{ci}_cfg"] = np.array([sr, tgt], np.int64)
    for ci, (sr, win, hop, n) in enumerate(RMS_CASES):
        s1, s2 = signal(300 + ci, n), signal(400 + ci, n)
        out[f"rms{ci}_s1"] = s1
        out[f"rms{ci}_s2"] = s2
        out[f"rms{ci}_cfg"] = np.array([sr, win, hop, 0.05], np.float64)
        out[f"rms{ci}_rms1"] = ews.frame_rms(s1, sr, win, hop)
        out[f"rms{ci}_rms2"] = ews.frame_rms(s2, sr, win, hop)
        out[f"rms{ci}_mask"] = ews.build_gt_overlap_mask(s1, s2, sr, win, hop, 0.05)
    np.savez_compressed(OUT, **out)
    print("wrote", OUT, os.path.getsize(OUT), "bytes")


if __name__ == "__main__":
    main()
