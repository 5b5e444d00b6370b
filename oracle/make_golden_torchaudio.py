"""Golden vectors for the MODEL FORWARD from an independent third-party implementation of Conv-TasNet that IS installed in
the build container: ``torchaudio.models.ConvTasNet`` (torchaudio 2.11).  asteroid 0.7.0 itself -- the package the reference
imports (/root/reference/src/osd/separation.py:40-46) -- is not vendored and not installable here, so this is the strongest
pin available for oracle/asteroid_restated.py: same paper, same layer graph, different authors.

    python oracle/make_golden_torchaudio.py      ->  tests/golden/torchaudio_convtasnet_golden.npz

What is stored: for each case the seed of ``oracle.asteroid_restated.make_model`` (weights are regenerated, not stored), the
input waveform and torchaudio's output.  The oracle's weights are copied into the torchaudio module by name
(``to_torchaudio``).  Two structural differences, handled outside the networks:
  * torchaudio's encoder / decoder convolutions carry ``padding = stride`` (one extra frame at either end over zero
    padding); asteroid's have none and ``pad_x_to_y`` right-pads the decoded signal.  Hence
        torchaudio(x) == asteroid(pad(x, stride, stride))[..., stride:-stride]     for len(x) % stride == 0;
  * torchaudio's last block has no residual conv; asteroid's has one whose output is never used.
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import asteroid_restated as ar  # noqa: E402

CASES = [  # (n_src, seed, samples, mask_act, kernel, stride)
    (3, 11, 8000, "sigmoid", 16, 8),
    (3, 12, 2400, "sigmoid", 16, 8),
    (2, 13, 4000, "relu", 16, 8),
    (3, 14, 4800, "sigmoid", 32, 16),
]


def to_torchaudio(m: "ar.ConvTasNet", n_src: int, mask_act: str, kernel: int, stride: int):
    from torchaudio.models import ConvTasNet
    assert stride * 2 == kernel
    t = ConvTasNet(num_sources=n_src, enc_kernel_size=kernel, enc_num_feats=512, msk_kernel_size=3, msk_num_feats=128,
                   msk_num_hidden_feats=512, msk_num_layers=8, msk_num_stacks=3, msk_activate=mask_act).eval()
    sd = m.state_dict()
    out = {"encoder.weight": sd["encoder.filterbank._filters"], "decoder.weight": sd["decoder.filterbank._filters"],
           "mask_generator.input_norm.weight": sd["masker.bottleneck.0.gamma"],
           "mask_generator.input_norm.bias": sd["masker.bottleneck.0.beta"],
           "mask_generator.input_conv.weight": sd["masker.bottleneck.1.weight"],
           "mask_generator.input_conv.bias": sd["masker.bottleneck.1.bias"],
           "mask_generator.output_prelu.weight": sd["masker.mask_net.0.weight"],
           "mask_generator.output_conv.weight": sd["masker.mask_net.1.weight"],
           "mask_generator.output_conv.bias": sd["masker.mask_net.1.bias"]}
    for i in range(24):
        a, b = f"masker.TCN.{i}.", f"mask_generator.conv_layers.{i}."
        for j, (wa, wb) in enumerate([("weight", "weight"), ("weight", "weight"), ("gamma", "weight"),
                                      ("weight", "weight"), ("weight", "weight"), ("gamma", "weight")]):
            out[f"{b}conv_layers.{j}.{wb}"] = sd[f"{a}shared_block.{j}.{wa}"]
        for j, (ba, bb) in {0: ("bias", "bias"), 2: ("beta", "bias"), 3: ("bias", "bias"), 5: ("beta", "bias")}.items():
            out[f"{b}conv_layers.{j}.{bb}"] = sd[f"{a}shared_block.{j}.{ba}"]
        out[f"{b}skip_out.weight"] = sd[f"{a}skip_conv.weight"]
        out[f"{b}skip_out.bias"] = sd[f"{a}skip_conv.bias"]
        if i < 23:  # torchaudio's last block has no residual conv (asteroid's is dead code)
            out[f"{b}res_out.weight"] = sd[f"{a}res_conv.weight"]
            out[f"{b}res_out.bias"] = sd[f"{a}res_conv.bias"]
    missing, unexpected = t.load_state_dict(out, strict=True)
    assert not missing and not unexpected
    return t


def main():
    import torchaudio
    arrays = {"torchaudio_version": np.array(torchaudio.__version__)}
    for ci, (n_src, seed, n, act, kernel, stride) in enumerate(CASES):
        m = ar.make_model(n_src, seed=seed, mask_act=act, kernel_size=kernel, stride=stride)
        t = to_torchaudio(m, n_src, act, kernel, stride)
        g = torch.Generator().manual_seed(1000 + seed)
        x = (0.1 * torch.randn(n, generator=g)).float()
        with torch.no_grad():
            y = t(x[None, None])[0].numpy()          # (n_src, n)
        arrays[f"case{ci}_meta"] = np.array([n_src, seed, n, kernel, stride, 1 if act == "relu" else 0], dtype=np.int64)
        arrays[f"case{ci}_x"] = x.numpy()
        arrays[f"case{ci}_y"] = y.astype(np.float32)
    path = os.path.join(ROOT, "tests", "golden", "torchaudio_convtasnet_golden.npz")
    np.savez_compressed(path, **arrays)
    print("wrote", path, {k: v.shape for k, v in arrays.items()})


if __name__ == "__main__":
    main()
