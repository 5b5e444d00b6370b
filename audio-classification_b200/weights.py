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


class PackedWeights:
    """Owns contiguous fp32 host copies of a state dict and the ``b200sep_weights`` struct that
    points into them (kept alive until ``b200sep_model_create`` has copied everything)."""

    def __init__(self, sd: Dict[str, Any], args: Dict[str, Any]):
        self.args = args
        self._keep = []
        n_layers = int(args["n_blocks"]) * int(args["n_repeats"])

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
