"""Checkpoint handling for the B200 separation backend.

Reads the same files the reference reads (``/root/reference/src/osd/separation.py:52-56``):
``torch.load`` of either a bare ``state_dict`` or a dict wrapping one under ``"state_dict"``; and
additionally honours Asteroid's ``model_args`` (``{model_name, state_dict, model_args}``, what
``BaseModel.from_pretrained`` consumes at ``separation.py:44-46``) so that e.g. the
16 kHz Libri3Mix model (kernel_size 32, stride 16, relu mask) runs with its real hyper-parameters.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Any, Dict, Optional, Tuple

import numpy as np

from . import _native as nat

DEFAULT_ARGS = dict(n_src=2, n_blocks=8, n_repeats=3, bn_chan=128, hid_chan=512, skip_chan=128,
                    conv_kernel_size=3, norm_type="gLN", mask_act="sigmoid", causal=False,
                    fb_name="free", kernel_size=16, n_filters=512, stride=8,
                    encoder_activation=None, sample_rate=8000)


def model_args_from_state_dict(sd: Dict[str, Any], n_src: Optional[int] = None) -> Dict[str, Any]:
    """Derive hyper-parameters from tensor shapes (used when a checkpoint has no ``model_args``)."""
    enc = sd["encoder.filterbank._filters"]
    n_filters, _, kernel_size = tuple(enc.shape)
    bn_chan = int(sd["masker.bottleneck.1.weight"].shape[0])
    hid_chan = int(sd["masker.TCN.0.shared_block.0.weight"].shape[0])
    skip_chan = int(sd["masker.TCN.0.skip_conv.weight"].shape[0])
    P = int(sd["masker.TCN.0.shared_block.3.weight"].shape[-1])
    n_layers = 1 + max(int(k.split(".")[2]) for k in sd if k.startswith("masker.TCN."))
    out = int(sd["masker.mask_net.1.weight"].shape[0])
    args = dict(DEFAULT_ARGS)
    args.update(n_filters=int(n_filters), kernel_size=int(kernel_size), stride=int(kernel_size) // 2,
                bn_chan=bn_chan, hid_chan=hid_chan, skip_chan=skip_chan, conv_kernel_size=P,
                n_src=out // int(n_filters))
    # n_blocks x n_repeats is not recoverable from shapes alone; assume the Asteroid default X=8
    if n_layers % 8 == 0:
        args.update(n_blocks=8, n_repeats=n_layers // 8)
    else:
        args.update(n_blocks=n_layers, n_repeats=1)
    if n_src is not None and args["n_src"] != n_src:
        raise RuntimeError(f"checkpoint separates {args['n_src']} sources, n_src={n_src} requested")
    return args


def load_checkpoint(path: str, n_src: Optional[int] = None) -> Tuple[Dict[str, Any], Dict[str, Any]]:
    """Returns ``(state_dict, model_args)``.  Missing file -> FileNotFoundError (as the reference,
    ``separation.py:124-129``)."""
    import torch
    if not os.path.isfile(path):
        raise FileNotFoundError(f"Separator checkpoint not found: {path}")
    # tensors and plain containers only (what Asteroid checkpoints hold); B200SEP_UNSAFE_LOAD=1 opts into full
    # unpickling for exotic files -- never the default: --sep-checkpoint may point at a downloaded file
    try:
        state = torch.load(path, map_location="cpu", weights_only=True)
    except Exception:
        if os.environ.get("B200SEP_UNSAFE_LOAD", "0") != "1":
            raise
        state = torch.load(path, map_location="cpu", weights_only=False)
    sd = state.get("state_dict", state)
    if isinstance(state, dict) and isinstance(state.get("model_args"), dict):
        args = dict(DEFAULT_ARGS)
        args.update(state["model_args"])
    else:
        args = model_args_from_state_dict(sd, n_src)
    return sd, args


def random_state_dict(n_src: int = 3, seed: int = 0, **overrides) -> Tuple[Dict[str, Any], Dict[str, Any]]:
    """Seeded random-init weights with Asteroid's key names and shapes (no checkpoint available
    offline -- BASELINE.json config 1 is quoted on random-init).  gLN gamma/beta and PReLU slopes
    are drawn away from their (1, 0, 0.25) defaults so the folded-norm paths are exercised."""
    import torch
    a = dict(DEFAULT_ARGS)
    a.update(overrides)
    a["n_src"] = n_src
    g = torch.Generator().manual_seed(int(seed))
    N, L, B, H, Sc, P = a["n_filters"], a["kernel_size"], a["bn_chan"], a["hid_chan"], a["skip_chan"], a["conv_kernel_size"]

    def uni(shape, fan_in):
        bound = 1.0 / float(np.sqrt(fan_in))
        return (torch.rand(shape, generator=g) * 2 - 1) * bound

    sd: Dict[str, Any] = {}
    std = float(np.sqrt(2.0 / (L + N * L)))
    sd["encoder.filterbank._filters"] = torch.randn((N, 1, L), generator=g) * std
    sd["decoder.filterbank._filters"] = torch.randn((N, 1, L), generator=g) * std
    sd["masker.bottleneck.0.gamma"] = 1.0 + 0.25 * torch.randn((N,), generator=g)
    sd["masker.bottleneck.0.beta"] = 0.1 * torch.randn((N,), generator=g)
    sd["masker.bottleneck.1.weight"] = uni((B, N, 1), N)
    sd["masker.bottleneck.1.bias"] = uni((B,), N)
    for i in range(a["n_blocks"] * a["n_repeats"]):
        p = f"masker.TCN.{i}."
        sd[p + "shared_block.0.weight"] = uni((H, B, 1), B)
        sd[p + "shared_block.0.bias"] = uni((H,), B)
        sd[p + "shared_block.1.weight"] = 0.05 + 0.4 * torch.rand((1,), generator=g)
        sd[p + "shared_block.2.gamma"] = 1.0 + 0.25 * torch.randn((H,), generator=g)
        sd[p + "shared_block.2.beta"] = 0.1 * torch.randn((H,), generator=g)
        sd[p + "shared_block.3.weight"] = uni((H, 1, P), P)
        sd[p + "shared_block.3.bias"] = uni((H,), P)
        sd[p + "shared_block.4.weight"] = 0.05 + 0.4 * torch.rand((1,), generator=g)
        sd[p + "shared_block.5.gamma"] = 1.0 + 0.25 * torch.randn((H,), generator=g)
        sd[p + "shared_block.5.beta"] = 0.1 * torch.randn((H,), generator=g)
        sd[p + "res_conv.weight"] = uni((B, H, 1), H)
        sd[p + "res_conv.bias"] = uni((B,), H)
        sd[p + "skip_conv.weight"] = uni((Sc, H, 1), H)
        sd[p + "skip_conv.bias"] = uni((Sc,), H)
    sd["masker.mask_net.0.weight"] = 0.05 + 0.4 * torch.rand((1,), generator=g)
    sd["masker.mask_net.1.weight"] = uni((n_src * N, Sc, 1), Sc)
    sd["masker.mask_net.1.bias"] = uni((n_src * N,), Sc)
    return sd, a


GLN_EPS = 1e-8  # asteroid.masknn.norms.GlobLN


def expected_shapes(args: Dict[str, Any]) -> Dict[str, Tuple[int, ...]]:
    """Every tensor of an Asteroid ConvTasNet state dict the backend reads and its shape."""
    N, L, B, H, Sc, P, S = (int(args[k]) for k in ("n_filters", "kernel_size", "bn_chan", "hid_chan", "skip_chan",
                                                    "conv_kernel_size", "n_src"))
    sh: Dict[str, Tuple[int, ...]] = {
        "encoder.filterbank._filters": (N, 1, L), "decoder.filterbank._filters": (N, 1, L),
        "masker.bottleneck.0.gamma": (N,), "masker.bottleneck.0.beta": (N,),
        "masker.bottleneck.1.weight": (B, N, 1), "masker.bottleneck.1.bias": (B,),
        "masker.mask_net.0.weight": (1,), "masker.mask_net.1.weight": (S * N, Sc, 1), "masker.mask_net.1.bias": (S * N,),
    }
    for i in range(int(args["n_blocks"]) * int(args["n_repeats"])):
        p = f"masker.TCN.{i}."
        sh.update({p + "shared_block.0.weight": (H, B, 1), p + "shared_block.0.bias": (H,), p + "shared_block.1.weight": (1,),
                   p + "shared_block.2.gamma": (H,), p + "shared_block.2.beta": (H,),
                   p + "shared_block.3.weight": (H, 1, P), p + "shared_block.3.bias": (H,), p + "shared_block.4.weight": (1,),
                   p + "shared_block.5.gamma": (H,), p + "shared_block.5.beta": (H,),
                   p + "res_conv.weight": (B, H, 1), p + "res_conv.bias": (B,),
                   p + "skip_conv.weight": (Sc, H, 1), p + "skip_conv.bias": (Sc,)})
    return sh


def validate_state_dict(sd: Dict[str, Any], args: Dict[str, Any]) -> None:
    for key, shape in expected_shapes(args).items():
        if key not in sd:
            raise RuntimeError(f"checkpoint is missing '{key}'")
        if tuple(sd[key].shape) != tuple(shape):
            raise RuntimeError(f"'{key}' has shape {tuple(sd[key].shape)}, expected {tuple(shape)}")


def _np64(t) -> np.ndarray:
    a = t.detach().cpu().numpy() if hasattr(t, "detach") else np.asarray(t)
    return np.asarray(a, dtype=np.float64)


def _pow2_exp(scale: float, lim: int = 40) -> int:
    """Exponent e such that 2^e is the power of two nearest to 1 / scale (0 for degenerate scales)."""
    if not np.isfinite(scale) or scale <= 0.0:
        return 0
    return int(np.clip(np.rint(-np.log2(scale)), -lim, lim))


def condition_state_dict(sd: Dict[str, Any], args: Dict[str, Any]):
    """Range conditioning: an EXACT power-of-two reparametrisation of a Conv-TasNet state dict that moves every
    tensor the fp16 tensor-core pipeline stores BEFORE a normalisation (encoder output, h1, h2, raw res / skip GEMM
    outputs, the residual and skip streams) to O(1) magnitude, whatever the scale of the checkpoint's weights.

    It uses two facts of the architecture: PReLU and the convolutions are positively homogeneous,
    so scaling a conv's weight AND bias by c scales its (post-PReLU) output by c; and the gLN that follows divides the
    scale out again, exactly so when its eps is scaled by c^2.  The residual stream, the skip sum and the encoder
    output / decoder pair are rescaled with the matching inverse on their consumers.  Powers of two make every
    product exact in fp32, so the conditioned model is the same function (the fp32 path stays >= 100 dB on it).

    Scales come from weight norms only (a white, unit-variance model of every normalised tensor), i.e. they are a
    property of the checkpoint, never of the audio: results stay independent of batch composition.

    Returns ``(conditioned state dict of float32 ndarrays, eps float32[1 + 2 n_layers], info dict)``.
    """
    n_layers = int(args["n_blocks"]) * int(args["n_repeats"])
    out: Dict[str, Any] = {k: np.array(_np64(sd[k]), dtype=np.float32) for k in expected_shapes(args)}
    t = "masker.TCN.{}."

    def scale(key: str, e: int) -> None:
        if e:
            out[key] = np.ldexp(out[key], e).astype(np.float32)

    # second moments under the white model (normalised tensor: zero mean, unit variance per element)
    def lin_m2(Wk: str, bk: str, gk: str, bek: str) -> np.ndarray:
        Wm, g, be, b = _np64(sd[Wk])[..., 0], _np64(sd[gk]), _np64(sd[bek]), _np64(sd[bk])
        return ((Wm * g[None, :]) ** 2).sum(1) + (Wm @ be + b) ** 2

    enc = _np64(sd["encoder.filterbank._filters"])[:, 0, :]
    a = _pow2_exp(float(np.sqrt((enc ** 2).sum(1).mean())) / 4.0)   # filter norm ~ 4: |tf| <= 16 for |wav| <= 1
    vx = lin_m2("masker.bottleneck.1.weight", "masker.bottleneck.1.bias", "masker.bottleneck.0.gamma",
                "masker.bottleneck.0.beta")
    vx_before = []
    vs = np.zeros_like(vx)
    for i in range(n_layers):
        p = t.format(i)
        vx_before.append(vx.copy())
        vx = vx + lin_m2(p + "res_conv.weight", p + "res_conv.bias", p + "shared_block.5.gamma", p + "shared_block.5.beta")
        vs = vs + lin_m2(p + "skip_conv.weight", p + "skip_conv.bias", p + "shared_block.5.gamma", p + "shared_block.5.beta")
    r = _pow2_exp(float(np.sqrt(vx.mean())))
    ts = _pow2_exp(float(np.sqrt(vs.mean())))
    eps = np.full(1 + 2 * n_layers, GLN_EPS, dtype=np.float64)
    ps, qs = [], []
    for i in range(n_layers):
        p = t.format(i)
        Win, bin_ = _np64(sd[p + "shared_block.0.weight"])[..., 0], _np64(sd[p + "shared_block.0.bias"])
        m2_h1 = (Win ** 2) @ vx_before[i] + bin_ ** 2
        pe = _pow2_exp(float(np.sqrt(m2_h1.mean())))
        g1, be1 = _np64(sd[p + "shared_block.2.gamma"]), _np64(sd[p + "shared_block.2.beta"])
        wd, bd = _np64(sd[p + "shared_block.3.weight"])[:, 0, :], _np64(sd[p + "shared_block.3.bias"])
        m2_h2 = g1 ** 2 * (wd ** 2).sum(1) + (be1 * wd.sum(1) + bd) ** 2
        qe = _pow2_exp(float(np.sqrt(m2_h2.mean())))
        ps.append(pe)
        qs.append(qe)
        scale(p + "shared_block.0.weight", pe - r)   # consumes x' = 2^r x, produces h1' = 2^pe h1
        scale(p + "shared_block.0.bias", pe)
        scale(p + "shared_block.3.weight", qe)       # consumes the normalised h1, produces h2' = 2^qe h2
        scale(p + "shared_block.3.bias", qe)
        scale(p + "res_conv.weight", r)
        scale(p + "res_conv.bias", r)
        scale(p + "skip_conv.weight", ts)
        scale(p + "skip_conv.bias", ts)
        eps[1 + 2 * i] = GLN_EPS * 4.0 ** pe
        eps[2 + 2 * i] = GLN_EPS * 4.0 ** qe
    scale("encoder.filterbank._filters", a)
    scale("decoder.filterbank._filters", -a)         # out = dec(mask * tf): undoes the encoder's 2^a
    eps[0] = GLN_EPS * 4.0 ** a
    scale("masker.bottleneck.1.weight", r)
    scale("masker.bottleneck.1.bias", r)
    scale("masker.mask_net.1.weight", -ts)           # consumes PReLU(2^ts skip_sum)
    info = dict(enc=a, x=r, skip=ts, h1=ps, h2=qs)
    return out, eps.astype(np.float32), info


class PackedWeights:
    """Owns contiguous fp32 host copies of a state dict and the ``b200sep_weights`` struct that
    points into them (kept alive until ``b200sep_model_create`` has copied everything).  ``condition=True`` (default;
    ``B200SEP_CONDITION=0`` disables) applies ``condition_state_dict`` first."""

    def __init__(self, sd: Dict[str, Any], args: Dict[str, Any], condition: Optional[bool] = None):
        self.args = args
        self._keep = []
        n_layers = int(args["n_blocks"]) * int(args["n_repeats"])
        if condition is None:
            condition = os.environ.get("B200SEP_CONDITION", "1") != "0"
        self.conditioning = None
        self.eps = np.full(1 + 2 * n_layers, GLN_EPS, dtype=np.float32)
        validate_state_dict(sd, args)
        if condition:
            sd, self.eps, self.conditioning = condition_state_dict(sd, args)

        def arr(key: str, shape) -> np.ndarray:
            if key not in sd:
                raise RuntimeError(f"checkpoint is missing '{key}'")
            t = sd[key]
            a = (t.detach().cpu().numpy() if hasattr(t, "detach") else np.asarray(t))
            a = np.ascontiguousarray(a, dtype=np.float32)
            if tuple(a.shape) != tuple(shape):
                raise RuntimeError(f"'{key}' has shape {tuple(a.shape)}, expected {tuple(shape)}")
            self._keep.append(a)
            return a

        def fp(a: np.ndarray):
            return a.ctypes.data_as(C.POINTER(C.c_float))

        def fpp(fmt: str, shape):
            ptrs = (C.POINTER(C.c_float) * n_layers)(*[fp(arr(fmt.format(i), shape)) for i in range(n_layers)])
            self._keep.append(ptrs)
            return C.cast(ptrs, C.POINTER(C.POINTER(C.c_float)))

        N, L, B, H, Sc, P, S = (int(args[k]) for k in ("n_filters", "kernel_size", "bn_chan", "hid_chan",
                                                        "skip_chan", "conv_kernel_size", "n_src"))
        w = nat.Weights()
        w.enc_filters = fp(arr("encoder.filterbank._filters", (N, 1, L)))
        w.dec_filters = fp(arr("decoder.filterbank._filters", (N, 1, L)))
        w.bn_gamma = fp(arr("masker.bottleneck.0.gamma", (N,)))
        w.bn_beta = fp(arr("masker.bottleneck.0.beta", (N,)))
        w.bn_weight = fp(arr("masker.bottleneck.1.weight", (B, N, 1)))
        w.bn_bias = fp(arr("masker.bottleneck.1.bias", (B,)))
        t = "masker.TCN.{}."
        w.in_weight = fpp(t + "shared_block.0.weight", (H, B, 1))
        w.in_bias = fpp(t + "shared_block.0.bias", (H,))
        w.prelu1 = fpp(t + "shared_block.1.weight", (1,))
        w.norm1_gamma = fpp(t + "shared_block.2.gamma", (H,))
        w.norm1_beta = fpp(t + "shared_block.2.beta", (H,))
        w.dw_weight = fpp(t + "shared_block.3.weight", (H, 1, P))
        w.dw_bias = fpp(t + "shared_block.3.bias", (H,))
        w.prelu2 = fpp(t + "shared_block.4.weight", (1,))
        w.norm2_gamma = fpp(t + "shared_block.5.gamma", (H,))
        w.norm2_beta = fpp(t + "shared_block.5.beta", (H,))
        w.res_weight = fpp(t + "res_conv.weight", (B, H, 1))
        w.res_bias = fpp(t + "res_conv.bias", (B,))
        w.skip_weight = fpp(t + "skip_conv.weight", (Sc, H, 1))
        w.skip_bias = fpp(t + "skip_conv.bias", (Sc,))
        w.mask_prelu = fp(arr("masker.mask_net.0.weight", (1,)))
        w.mask_weight = fp(arr("masker.mask_net.1.weight", (S * N, Sc, 1)))
        w.mask_bias = fp(arr("masker.mask_net.1.bias", (S * N,)))
        self.eps = np.ascontiguousarray(self.eps, dtype=np.float32)
        w.gln_eps = fp(self.eps)
        self.struct = w

    def config(self, precision: str) -> "nat.Config":
        a = self.args
        if a.get("norm_type", "gLN") != "gLN" or a.get("causal", False):
            raise RuntimeError("only the non-causal gLN Conv-TasNet is supported")
        if a.get("fb_name", "free") != "free":
            raise RuntimeError("only the 'free' filterbank is supported")
        if a.get("encoder_activation") not in (None, "linear"):
            raise RuntimeError("encoder_activation must be None")
        if a.get("mask_act", "sigmoid") not in nat.MASK_ACTS:
            raise RuntimeError(f"unsupported mask_act {a.get('mask_act')!r}")
        if precision not in nat.PRECISIONS:
            raise RuntimeError(f"unknown precision {precision!r}; choose from {sorted(nat.PRECISIONS)}")
        c = nat.Config()
        c.n_src, c.n_filters, c.kernel_size, c.stride = int(a["n_src"]), int(a["n_filters"]), int(a["kernel_size"]), int(a["stride"])
        c.bn_chan, c.hid_chan, c.skip_chan = int(a["bn_chan"]), int(a["hid_chan"]), int(a["skip_chan"])
        c.conv_kernel_size, c.n_blocks, c.n_repeats = int(a["conv_kernel_size"]), int(a["n_blocks"]), int(a["n_repeats"])
        c.mask_act = nat.MASK_ACTS[a.get("mask_act", "sigmoid")]
        c.precision = nat.PRECISIONS[precision]
        return c
