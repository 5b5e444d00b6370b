"""Device versions of the code either side of the separation path, through the C ABI.

* ``resample_batch``          -- ``Separator._ensure_sr`` (``/root/reference/src/osd/separation.py:91-103``, method
                                 ``"linear"``, bit-identical for float32 input) and ``torchaudio.functional.resample``
                                 as the reference's loaders call it (``scripts/osd/overlap3_core.py:136-143``,
                                 ``scripts/osd/evaluate_with_sources.py:163-178``; method ``"sinc"``)
* ``frame_rms`` / ``build_gt_overlap_mask`` -- same names and arguments as
                                 ``scripts/osd/evaluate_with_sources.py:181-197, 226-241`` (+ ``*_batch`` forms)
* ``slice_refs``              -- the reference-slicing rule of the evaluation call site
                                 (``scripts/osd/overlap3_core.py:716-722``: ``snp[int(s * sr):int(e * sr)]``)

There is no CPU fallback: every function raises ``RuntimeError`` when the native library or a CUDA device is missing.
"""
from __future__ import annotations

import ctypes as C
from typing import List, Sequence, Tuple

import numpy as np

from . import _native as nat

METHODS = {"linear": 0, "sinc": 1}


def _torch_cuda(device: int):
    import torch
    if not torch.cuda.is_available():
        raise RuntimeError("b200sep.aux needs a CUDA (sm_100a) device; there is no CPU fallback")
    return torch, torch.device("cuda", device)


def _pack_f32(signals: Sequence[np.ndarray]) -> Tuple[np.ndarray, List[int]]:
    arrs = [np.ascontiguousarray(np.asarray(s).reshape(-1), dtype=np.float32) for s in signals]
    lens = [int(a.shape[0]) for a in arrs]
    return (np.concatenate(arrs) if arrs else np.zeros(0, np.float32)), lens


def resample_out_lens(lens: Sequence[int], sr_in: int, sr_out: int, method: str = "linear") -> List[int]:
    """Host-only: output lengths (the reference's rounding and pass-through rules)."""
    n = len(lens)
    if n == 0:
        return []
    out = (C.c_int32 * n)()
    nat.check(nat.lib().b200sep_resample_out_lens(nat.i32_array(lens), n, int(sr_in), int(sr_out), METHODS[method], out),
              "b200sep_resample_out_lens")
    return [int(v) for v in out]


def resample_device(d_in, lens: Sequence[int], sr_in: int, sr_out: int, method: str = "linear"):
    """``d_in``: packed 1-D float32 CUDA tensor; returns (packed float32 CUDA tensor, out_lens).  Asynchronous."""
    import torch
    olens = resample_out_lens(lens, sr_in, sr_out, method)
    d_out = torch.empty(max(int(sum(olens)), 1), dtype=torch.float32, device=d_in.device)
    with torch.cuda.device(d_in.device):
        st = torch.cuda.current_stream(d_in.device).cuda_stream
        nat.check(nat.lib().b200sep_resample_batch(C.c_void_p(d_in.data_ptr()), nat.i32_array(lens), len(lens), int(sr_in),
                                                   int(sr_out), METHODS[method], C.c_void_p(d_out.data_ptr()),
                                                   C.c_void_p(st)), "b200sep_resample_batch")
    return d_out, olens


def resample_batch(signals: Sequence[np.ndarray], sr_in: int, sr_out: int, method: str = "linear",
                   device: int = 0) -> List[np.ndarray]:
    """numpy in -> numpy out.  ``linear`` follows ``_ensure_sr`` (same rate / <= 1 sample / target <= 1 sample return
    the input unchanged); ``sinc`` follows ``AF.resample`` behind the loaders' ``numel() > 1`` guard."""
    torch, dev = _torch_cuda(device)
    signals = list(signals)
    if len(signals) == 0:
        return []
    if sr_in == sr_out:
        return signals
    lens = [int(np.asarray(s).shape[-1]) for s in signals]
    olens = resample_out_lens(lens, sr_in, sr_out, method)
    # pass-through signals (<= 1 sample; linear: target length <= 1) are returned as they came
    todo = [i for i in range(len(signals)) if lens[i] > 1 and
            not (method == "linear" and int(round(lens[i] * sr_out / sr_in)) <= 1)]
    res: List[np.ndarray] = list(signals)
    if not todo:
        return res
    flat, tl = _pack_f32([signals[i] for i in todo])
    d_in = torch.from_numpy(flat).to(dev)
    d_out, ol = resample_device(d_in, tl, sr_in, sr_out, method)
    h = d_out.cpu().numpy()
    o = 0
    for i, n in zip(todo, ol):
        res[i] = h[o:o + n]
        o += n
    return res


def frame_rms_out_lens(lens: Sequence[int], win_s: int, hop_s: int) -> List[int]:
    n = len(lens)
    if n == 0:
        return []
    out = (C.c_int32 * n)()
    rc = nat.lib().b200sep_frame_rms_out_lens(nat.i32_array(lens), n, int(win_s), int(hop_s), out)
    if rc != 0:
        raise ValueError(nat.last_error())  # "win too small" / "hop too small", as the reference raises
    return [int(v) for v in out]


def frame_rms_batch(signals: Sequence[np.ndarray], sr: int, win: float, hop: float, device: int = 0) -> List[np.ndarray]:
    torch, dev = _torch_cuda(device)
    win_s, hop_s = int(win * sr), int(hop * sr)
    if win_s <= 0:
        raise ValueError("win too small")
    if hop_s <= 0:
        raise ValueError("hop too small")
    if len(signals) == 0:
        return []
    flat, lens = _pack_f32(signals)
    fl = frame_rms_out_lens(lens, win_s, hop_s)
    d_in = torch.from_numpy(flat).to(dev) if flat.size else torch.zeros(1, dtype=torch.float32, device=dev)
    d_rms = torch.empty(max(sum(fl), 1), dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        st = torch.cuda.current_stream(dev).cuda_stream
        nat.check(nat.lib().b200sep_frame_rms_batch(C.c_void_p(d_in.data_ptr()), nat.i32_array(lens), len(lens), win_s, hop_s,
                                                    C.c_void_p(d_rms.data_ptr()), C.c_void_p(st)), "b200sep_frame_rms_batch")
    h = d_rms.cpu().numpy()
    out, o = [], 0
    for n in fl:
        out.append(h[o:o + n])
        o += n
    return out


def frame_rms(wav: np.ndarray, sr: int, win: float, hop: float) -> np.ndarray:
    """evaluate_with_sources.py:181-197."""
    return frame_rms_batch([wav], sr, win, hop)[0]


def build_gt_overlap_mask_batch(pairs: Sequence[Tuple[np.ndarray, np.ndarray]], sr: int, win: float, hop: float,
                                thr_ratio: float, device: int = 0) -> List[np.ndarray]:
    """One boolean frame mask per (s1, s2) pair; RMS tracks and masks never leave the device in between."""
    torch, dev = _torch_cuda(device)
    win_s, hop_s = int(win * sr), int(hop * sr)
    if win_s <= 0:
        raise ValueError("win too small")
    if hop_s <= 0:
        raise ValueError("hop too small")
    if len(pairs) == 0:
        return []
    f1, l1 = _pack_f32([p[0] for p in pairs])
    f2, l2 = _pack_f32([p[1] for p in pairs])
    n1, n2 = frame_rms_out_lens(l1, win_s, hop_s), frame_rms_out_lens(l2, win_s, hop_s)
    if n1 != n2:
        raise ValueError("s1 and s2 of a pair must yield the same number of frames")  # numpy would fail to broadcast
    with torch.cuda.device(dev):
        st = torch.cuda.current_stream(dev).cuda_stream
        tot = max(sum(n1), 1)
        d_r = []
        for flat, lens in ((f1, l1), (f2, l2)):
            d_in = torch.from_numpy(flat).to(dev) if flat.size else torch.zeros(1, dtype=torch.float32, device=dev)
            d_rms = torch.empty(tot, dtype=torch.float32, device=dev)
            nat.check(nat.lib().b200sep_frame_rms_batch(C.c_void_p(d_in.data_ptr()), nat.i32_array(lens), len(lens), win_s,
                                                        hop_s, C.c_void_p(d_rms.data_ptr()), C.c_void_p(st)),
                      "b200sep_frame_rms_batch")
            d_r.append((d_in, d_rms))
        d_mask = torch.empty(tot, dtype=torch.uint8, device=dev)
        nat.check(nat.lib().b200sep_gt_overlap_mask_batch(C.c_void_p(d_r[0][1].data_ptr()), C.c_void_p(d_r[1][1].data_ptr()),
                                                          nat.i32_array(n1), len(n1), C.c_float(float(thr_ratio)),
                                                          C.c_void_p(d_mask.data_ptr()), C.c_void_p(st)),
                  "b200sep_gt_overlap_mask_batch")
        h = d_mask.cpu().numpy().astype(bool)
    out, o = [], 0
    for n in n1:
        out.append(h[o:o + n])
        o += n
    return out


def build_gt_overlap_mask(s1: np.ndarray, s2: np.ndarray, sr: int, win: float, hop: float, thr_ratio: float) -> np.ndarray:
    """evaluate_with_sources.py:226-241."""
    return build_gt_overlap_mask_batch([(s1, s2)], sr, win, hop, thr_ratio)[0]


def slice_refs(refs: Sequence[np.ndarray], s: float, e: float, sr: int) -> List[np.ndarray]:
    """overlap3_core.py:716-722: every reference cut to the overlap segment with the mixture's own index rule."""
    s_i, e_i = int(s * sr), int(e * sr)
    return [np.asarray(r)[s_i:e_i] for r in refs]
