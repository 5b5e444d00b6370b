"""CPU tests of the oracle restatement of asteroid ConvTasNet."""
import hashlib
import os

import numpy as np
import pytest
import torch

from oracle import asteroid_restated as ar
from oracle import pit_restated as pr
from oracle import shim


def test_param_counts_match_asteroid_accounting():
    # SURVEY.md A.4
    assert ar.count_params(ar.ConvTasNet(n_src=3)) == 5_116_593
    assert ar.count_params(ar.ConvTasNet(n_src=2)) == 5_050_545


def test_state_dict_keys_and_shapes():
    sd = ar.ConvTasNet(n_src=3).state_dict()
    assert sd["encoder.filterbank._filters"].shape == (512, 1, 16)
    assert sd["decoder.filterbank._filters"].shape == (512, 1, 16)
    assert sd["masker.bottleneck.0.gamma"].shape == (512,)
    assert sd["masker.bottleneck.1.weight"].shape == (128, 512, 1)
    for i in (0, 23):
        p = f"masker.TCN.{i}."
        assert sd[p + "shared_block.0.weight"].shape == (512, 128, 1)
        assert sd[p + "shared_block.1.weight"].shape == (1,)
        assert sd[p + "shared_block.2.gamma"].shape == (512,)
        assert sd[p + "shared_block.3.weight"].shape == (512, 1, 3)
        assert sd[p + "shared_block.5.beta"].shape == (512,)
        assert sd[p + "res_conv.weight"].shape == (128, 512, 1)
        assert sd[p + "skip_conv.bias"].shape == (128,)
    assert sd["masker.mask_net.1.weight"].shape == (1536, 128, 1)
    assert len([k for k in sd if k.startswith("masker.TCN.")]) == 24 * 14


@pytest.mark.parametrize("T", [16, 23, 24, 1600, 4321])
def test_output_shape_and_zero_tail(T):
    m = ar.make_model(3, seed=3, n_blocks=2, n_repeats=1)
    x = torch.randn(1, T) * 0.05
    with torch.no_grad():
        y = m(x)
    assert y.shape == (1, 3, T)
    tail = (T - 16) % 8
    if tail:
        assert torch.all(y[..., -tail:] == 0)
        assert torch.any(y[..., : T - tail] != 0)


def test_fp64_self_consistency():
    m = ar.make_model(3, seed=1, n_blocks=4, n_repeats=2)
    x = torch.randn(1, 4000) * 0.05
    with torch.no_grad():
        y32 = m(x)[0].numpy()
        y64 = m.double()(x.double())[0].numpy()
    for s in range(3):
        assert pr.si_sdr_f64(y64[s], y32[s]) > 100.0


def test_block_matches_torchaudio_convblock():
    """Independent implementation of the same paper: torchaudio ConvBlock == Conv1DBlock."""
    ta = pytest.importorskip("torchaudio.models.conv_tasnet")
    torch.manual_seed(0)
    blk = ar.Conv1DBlock(128, 512, 128, 3, padding=4, dilation=4).eval()
    with torch.no_grad():
        for n, p in blk.named_parameters():
            if n.endswith("gamma"):
                p.copy_(1 + 0.2 * torch.randn_like(p))
            if n.endswith("beta"):
                p.copy_(0.1 * torch.randn_like(p))
    tb = ta.ConvBlock(io_channels=128, hidden_channels=512, kernel_size=3, padding=4, dilation=4, no_residual=False).eval()
    sb = blk.shared_block
    with torch.no_grad():
        tb.conv_layers[0].weight.copy_(sb[0].weight); tb.conv_layers[0].bias.copy_(sb[0].bias)
        tb.conv_layers[1].weight.copy_(sb[1].weight)
        tb.conv_layers[2].weight.copy_(sb[2].gamma); tb.conv_layers[2].bias.copy_(sb[2].beta)
        tb.conv_layers[3].weight.copy_(sb[3].weight); tb.conv_layers[3].bias.copy_(sb[3].bias)
        tb.conv_layers[4].weight.copy_(sb[4].weight)
        tb.conv_layers[5].weight.copy_(sb[5].gamma); tb.conv_layers[5].bias.copy_(sb[5].beta)
        tb.res_out.weight.copy_(blk.res_conv.weight); tb.res_out.bias.copy_(blk.res_conv.bias)
        tb.skip_out.weight.copy_(blk.skip_conv.weight); tb.skip_out.bias.copy_(blk.skip_conv.bias)
    assert tb.conv_layers[2].eps == 1e-8
    x = torch.randn(1, 128, 300)
    with torch.no_grad():
        r1, s1 = blk(x)
        r2, s2 = tb(x)
    assert torch.allclose(r1, r2, atol=2e-5, rtol=1e-4)
    assert torch.allclose(s1, s2, atol=2e-5, rtol=1e-4)


def test_whole_model_matches_torchaudio_convtasnet_golden(golden_dir):
    """PIN of the forward oracle: torchaudio.models.ConvTasNet -- an independent, installed implementation of the same
    network -- run on the oracle's weights (oracle/make_golden_torchaudio.py, outputs committed under tests/golden/).
    torchaudio(x) == asteroid(pad(x, stride, stride))[stride:-stride] (its encoder / decoder carry padding = stride)."""
    g = np.load(os.path.join(golden_dir, "torchaudio_convtasnet_golden.npz"))
    n_cases = sum(1 for k in g.files if k.endswith("_meta"))
    assert n_cases >= 4
    for ci in range(n_cases):
        n_src, seed, n, kernel, stride, relu = [int(v) for v in g[f"case{ci}_meta"]]
        m = ar.make_model(n_src, seed=seed, mask_act="relu" if relu else "sigmoid", kernel_size=kernel, stride=stride)
        x = torch.from_numpy(g[f"case{ci}_x"])
        with torch.no_grad():
            y = m(torch.nn.functional.pad(x, (stride, stride))[None])[0].numpy()[:, stride:-stride]
        ref = g[f"case{ci}_y"]
        assert y.shape == ref.shape == (n_src, n)
        worst = min(pr.si_sdr_f64(ref[s], y[s]) for s in range(n_src))
        assert worst >= 100.0, (ci, worst)


def test_torchaudio_golden_is_reproducible():
    """The committed golden equals what the generating script produces from the installed torchaudio (skipped where
    torchaudio is absent, e.g. a trimmed GPU box)."""
    pytest.importorskip("torchaudio.models")
    from oracle import make_golden_torchaudio as mg
    n_src, seed, n, act, kernel, stride = mg.CASES[1]
    m = ar.make_model(n_src, seed=seed, mask_act=act, kernel_size=kernel, stride=stride)
    t = mg.to_torchaudio(m, n_src, act, kernel, stride)
    gold = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "torchaudio_convtasnet_golden.npz"))
    x = torch.from_numpy(gold["case1_x"])
    with torch.no_grad():
        y = t(x[None, None])[0].numpy()
    assert min(pr.si_sdr_f64(gold["case1_y"][s], y[s]) for s in range(n_src)) >= 100.0


def test_checkpoint_roundtrip_from_pretrained(tmp_path):
    m = ar.make_model(2, seed=5, n_blocks=2, n_repeats=1, mask_act="relu", kernel_size=32, stride=16)
    path = str(tmp_path / "pytorch_model.bin")
    torch.save(m.serialize(), path)
    m2 = ar.BaseModel.from_pretrained(path).eval()
    assert m2.masker.mask_act == "relu" and m2.encoder.filterbank.kernel_size == 32
    x = torch.randn(1, 3000) * 0.1
    with torch.no_grad():
        assert torch.equal(m(x), m2(x))


def _digest(sd):
    h = hashlib.sha256()
    for k in sorted(sd):
        h.update(k.encode())
        h.update(np.ascontiguousarray(sd[k].detach().cpu().numpy(), dtype=np.float32).tobytes())
    return h.hexdigest()


def test_oracle_reproduces_separator_golden(golden_dir):
    """tests/golden/separator_golden.npz was produced by the reference's own Separator wrapper
    (src/osd/separation.py, unmodified) on top of this restatement; the oracle called directly on the
    same arrays and weights must reproduce it (>= 100 dB; fp32 summation order may differ with the
    host's thread count)."""
    from b200sep import weights
    g = np.load(os.path.join(golden_dir, "separator_golden.npz"))
    for n_src in (3, 2):
        sd, args = weights.random_state_dict(n_src, seed=0)
        if _digest(sd) != str(g[f"s{n_src}_digest"]):
            pytest.skip("torch RNG stream differs from the one that generated the golden weights")
        m = ar.ConvTasNet(**args).eval()
        m.load_state_dict(sd)
        for i in range(int(g[f"s{n_src}_n"])):
            mix, ref = g[f"s{n_src}_mix{i}"], g[f"s{n_src}_out{i}"]
            with torch.no_grad():
                out = m(torch.from_numpy(mix)[None])[0].numpy()
            assert out.shape == ref.shape
            for s in range(n_src):
                assert pr.si_sdr_f64(ref[s], out[s]) > 100.0


@pytest.mark.skipif(not shim.reference_available(), reason="/root/reference only exists in the build container")
def test_reference_separator_runs_unmodified_through_shim():
    from b200sep import weights
    Sep = shim.import_reference_separator()
    sd, args = weights.random_state_dict(3, seed=0, n_blocks=2, n_repeats=1)

    def factory():
        m = ar.ConvTasNet(**args)
        m.load_state_dict(sd)
        return m

    ar.BaseModel.register_pretrained("JorisCos/ConvTasNet_Libri3Mix_sepclean_16k", factory)
    sep = Sep(device="cpu", n_src=3, sample_rate=8000)
    x = (np.random.default_rng(0).standard_normal(2000) * 0.05)
    out = sep.separate(x, 8000)               # float64 in -> float32 out, same length
    assert len(out) == 3 and all(o.dtype == np.float32 and o.shape == (2000,) for o in out)
    out16 = sep.separate(x.astype(np.float32), 4000)   # sr != sample_rate -> linear resample x2
    assert out16[0].shape == (4000,)
