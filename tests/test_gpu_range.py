"""Dynamic-range stress of the fp16 tensor-core pipeline (VERDICT r01 "what's weak" #1).

All earlier parity evidence used random-init weights whose pre-gLN activations are O(1).  A trained checkpoint is
free to put any scale on the convolutions in front of a GlobLN -- the norm removes it, so the fp32 oracle is (almost)
invariant -- while the fp16 storage of h1 / h2 / G / the residual stream is not.  These tests scale exactly those
weights by s in {1e-4 .. 1e6} and draw gamma / beta / alpha from wide distributions, and hold the CUDA path (with the
load-time range conditioning of ``b200sep.weights.condition_state_dict``) to the 40 dB gate of BASELINE.json's
north_star against the oracle's fp32 output; they also show that without the conditioning the same weights overflow
fp16 and the engine re-runs the batch on its fp32 twin instead of returning NaNs.
"""
import warnings

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import asteroid_restated as ar  # noqa: E402
from oracle import pit_restated as pr  # noqa: E402

GATE_DB = 40.0   # BASELINE.json north_star: >= 40 dB SI-SDR relative to the reference's fp32 outputs


def _scaled(sd, s):
    out = {k: v.clone() for k, v in sd.items()}
    for k in out:
        if k.endswith(("shared_block.0.weight", "shared_block.0.bias", "shared_block.3.weight", "shared_block.3.bias")):
            out[k] = out[k] * s
    out["encoder.filterbank._filters"] = out["encoder.filterbank._filters"] * s
    return out


def _oracle(sd, args):
    m = ar.ConvTasNet(**args).eval()
    m.load_state_dict(sd)
    return m


def _worst_db(model, eng, mixes):
    outs = eng.separate_batch(mixes)
    worst = 1e9
    for m, o in zip(mixes, outs):
        with torch.no_grad():
            ref = model(torch.from_numpy(m)[None])[0].numpy()
        for s_ in range(ref.shape[0]):
            worst = min(worst, pr.si_sdr_f64(ref[s_], o[s_]))
    return worst


@pytest.fixture(scope="module")
def mixes():
    from b200sep import synth
    # configs[0] (one 4 s mixture) + a stratified slice of the configs[1] batch (seed 2 lengths, U[0.2, 4] s)
    lens = synth.segment_lengths(64, 2)
    pick = [32000] + [lens[i] for i in (0, 9, 23, 41, 63)]
    return [synth.make_mixture(700 + i, int(n), 3)[0] for i, n in enumerate(pick)]


@pytest.mark.parametrize("s", [1e-2, 1e2, 1e3, 1e4, 1e6])
def test_scaled_weights_stay_within_gate(s, mixes):
    import b200sep
    from b200sep import weights
    sd0, args = weights.random_state_dict(3, 0)
    sd = _scaled(sd0, s)
    model = _oracle(sd, args)
    eng = b200sep.SeparationEngine(sd, args, 0, "fp16")
    with warnings.catch_warnings():
        warnings.simplefilter("error")          # a fp32 re-run would warn: the fp16 pipeline itself must hold
        worst = _worst_db(model, eng, mixes)
    assert eng.fallback_reruns == 0
    print(f"[range] s={s:g}: worst SI-SDR vs oracle fp32 = {worst:.1f} dB; conditioning {eng.conditioning}")
    assert worst >= GATE_DB, (s, worst)
    eng.close()


@pytest.mark.parametrize("s", [1e-2, 1e3])
def test_scaled_residual_stream_stays_within_gate(s, mixes):
    """The residual stream is never normalised: scaling the bottleneck conv and every residual conv by s scales x, and with
    it the pre-gLN1 activations h1 = W_in x (stored in fp16, summed for the gLN1 statistics), by s.  The conditioning's
    white model of x has to take that scale out again (no fp32 re-run, >= 40 dB)."""
    import b200sep
    from b200sep import weights
    sd0, args = weights.random_state_dict(3, 0)
    sd = {k: v.clone() for k, v in sd0.items()}
    for k in sd:
        if k.startswith("masker.bottleneck.1.") or ".res_conv." in k:
            sd[k] = sd[k] * s
    model = _oracle(sd, args)
    eng = b200sep.SeparationEngine(sd, args, 0, "fp16")
    with warnings.catch_warnings():
        warnings.simplefilter("error")
        worst = _worst_db(model, eng, mixes[:4])
    assert eng.fallback_reruns == 0
    print(f"[range] residual stream x {s:g}: worst SI-SDR vs oracle fp32 = {worst:.1f} dB; conditioning {eng.conditioning}")
    assert worst >= GATE_DB, (s, worst)
    eng.close()


def test_eps_dominated_scale_is_detected_and_rerun_in_fp32(mixes):
    """s = 1e-4 on the encoder filters puts the encoder output's variance 1e4 below the bottleneck gLN's eps = 1e-8:
    the norm then divides by sqrt(eps), the signal reaches the TCN as a ripple ~250 x smaller than the biases, and fp16
    activation storage keeps 3 bits of it (4.7 dB measured).  The fp32 oracle handles it, so the half-precision pipeline
    must notice (per-call device flag bit 2) and the engine must answer from its fp32 twin."""
    import b200sep
    from b200sep import weights
    sd0, args = weights.random_state_dict(3, 0)
    sd = _scaled(sd0, 1e-4)
    model = _oracle(sd, args)
    eng = b200sep.SeparationEngine(sd, args, 0, "fp16")
    with pytest.warns(RuntimeWarning, match="far below its eps"):
        worst = _worst_db(model, eng, mixes[:3])
    assert eng.fallback_reruns == 1
    print(f"[range] s=1e-4 (eps-dominated): worst SI-SDR after the fp32 re-run = {worst:.1f} dB")
    assert worst >= GATE_DB, worst
    eng.close()


def _wide_affine(seed):
    """gamma log-normal over ~3 decades with random signs, beta ~ N(0, 1), PReLU slopes in [-0.3, 1.5] (outside [0, 1]
    the kernels take their general PReLU path)."""
    from b200sep import weights
    sd, args = weights.random_state_dict(3, seed)
    g = torch.Generator().manual_seed(seed + 31)
    for k in list(sd):
        if k.endswith(".gamma"):
            mag = torch.exp(1.2 * torch.randn(sd[k].shape, generator=g))
            sign = torch.where(torch.rand(sd[k].shape, generator=g) < 0.2, -1.0, 1.0)
            sd[k] = mag * sign
        elif k.endswith(".beta"):
            sd[k] = torch.randn(sd[k].shape, generator=g)
        elif sd[k].numel() == 1:
            sd[k] = -0.3 + 1.8 * torch.rand((1,), generator=g)
    return sd, args


@pytest.mark.parametrize("seed", [1, 2])
def test_wide_gamma_beta_alpha(seed, mixes):
    import b200sep
    sd, args = _wide_affine(seed)
    model = _oracle(sd, args)
    eng = b200sep.SeparationEngine(sd, args, 0, "fp16")
    worst = _worst_db(model, eng, mixes[:4])
    print(f"[range] wide affine seed {seed}: worst SI-SDR vs oracle fp32 = {worst:.1f} dB, fp32 re-runs {eng.fallback_reruns}")
    assert worst >= GATE_DB, worst
    eng.close()


def test_fp16_overflow_reruns_in_fp32(mixes, monkeypatch):
    """Without the conditioning, weights scaled by 1e6 push h1 / G out of the fp16 range (1e4 still fits: fp16 is a
    floating-point format).  The overflow is caught where it first shows -- the next gLN's variance is non-finite
    (per-call device flag bit 1; the output alone would not tell: the sigmoid turns +-inf scores into clean 0 / 1
    masks) -- and the engine repeats the batch on the fp32 pipeline: the caller gets oracle-grade waveforms plus a
    RuntimeWarning, never garbage."""
    import b200sep
    from b200sep import weights
    monkeypatch.setenv("B200SEP_CONDITION", "0")
    sd0, args = weights.random_state_dict(3, 0)
    sd = _scaled(sd0, 1e6)
    model = _oracle(sd, args)
    eng = b200sep.SeparationEngine(sd, args, 0, "fp16")
    assert eng.conditioning is None
    with pytest.warns(RuntimeWarning, match="re-run on the fp32 pipeline"):
        worst = _worst_db(model, eng, mixes[:2])
    print(f"[range] unconditioned s=1e6: worst SI-SDR after the fp32 re-run = {worst:.1f} dB")
    assert eng.fallback_reruns == 1
    assert worst >= 100.0, worst   # the fp32 pipeline's own gate
    eng.close()


def test_non_finite_input_is_reported(mixes):
    import b200sep
    from b200sep import weights
    sd, args = weights.random_state_dict(3, 0)
    eng = b200sep.SeparationEngine(sd, args, 0, "fp16")
    bad = mixes[1].copy()
    bad[100] = np.nan
    with pytest.warns(RuntimeWarning):
        out = eng.separate_batch([bad, mixes[2]])
    assert not np.all(np.isfinite(np.stack(out[0])))
    assert np.all(np.isfinite(np.stack(out[1])))      # the neighbour segment is untouched (independent units)
    eng.close()


def test_two_devices_in_one_process(mixes):
    """Per-device state (function attributes, staging rings, flags): engines on cuda:0 and cuda:1 of ONE process give
    bit-identical results.  Needs two GPUs (`gpurun --gpus 2`)."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import b200sep
    from b200sep import weights
    sd, args = weights.random_state_dict(3, 0)
    e0 = b200sep.SeparationEngine(sd, args, 0, "fp16")
    e1 = b200sep.SeparationEngine(sd, args, 1, "fp16")
    o1 = e1.separate_batch(mixes[:3])
    o0 = e0.separate_batch(mixes[:3])
    for a, b in zip(o0, o1):
        for x, y in zip(a, b):
            assert np.array_equal(x, y)
    res0 = b200sep.pit_sisdr_batch(mixes[:3], [[m, m, m] for m in mixes[:3]], o0, "overlap3")
    assert np.all(np.isfinite(res0.sisdr))
    e0.close()
    e1.close()
