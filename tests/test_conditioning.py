"""CPU tests of the load-time range conditioning (``b200sep.weights.condition_state_dict``): it must be an EXACT
reparametrisation -- the conditioned weights with the returned per-gLN eps values are the same function as the
original checkpoint -- and it must absorb any scale a checkpoint puts in front of a GlobLN."""
import numpy as np
import pytest
import torch

from oracle import asteroid_restated as ar


def _run(sd, args, eps, x):
    m = ar.ConvTasNet(**args).double().eval()
    m.load_state_dict({k: (v if torch.is_tensor(v) else torch.from_numpy(np.asarray(v))).double() for k, v in sd.items()})
    if eps is not None:
        ar.set_gln_eps(m, np.asarray(eps, dtype=np.float64))
    with torch.no_grad():
        return m(x).numpy()


@pytest.mark.parametrize("s", [1.0, 1e-4, 1e3])
def test_conditioning_is_an_exact_reparametrisation(s):
    from b200sep import weights
    sd, args = weights.random_state_dict(3, 3, n_blocks=3, n_repeats=2)
    sd = {k: v.clone() for k, v in sd.items()}
    for k in sd:
        if k.endswith(("shared_block.0.weight", "shared_block.0.bias", "shared_block.3.weight", "shared_block.3.bias")):
            sd[k] = sd[k] * s
    sd["encoder.filterbank._filters"] = sd["encoder.filterbank._filters"] * s
    sd2, eps, info = weights.condition_state_dict(sd, args)
    assert eps.shape == (1 + 2 * 6,) and eps.dtype == np.float32
    x = 0.05 * torch.randn(2, 4000, dtype=torch.float64, generator=torch.Generator().manual_seed(0))
    y = _run(sd, args, None, x)
    y2 = _run(sd2, args, eps, x)
    assert np.max(np.abs(y - y2)) <= 1e-9 * np.max(np.abs(y))
    # every multiplier is a power of two (exact in fp32) and eps follows its gLN's input scale squared
    for k in sd2:
        r = sd2[k].astype(np.float64).ravel() / np.where(sd[k].numpy().ravel() == 0, 1.0, sd[k].double().numpy().ravel())
        r = r[np.asarray(sd[k]).ravel() != 0]
        if r.size:
            e = np.log2(r[0])
            assert abs(e - round(e)) < 1e-12 and np.all(r == r[0]), k
    assert eps[0] == np.float32(1e-8 * 4.0 ** info["enc"])
    assert eps[1] == np.float32(1e-8 * 4.0 ** info["h1"][0]) and eps[2] == np.float32(1e-8 * 4.0 ** info["h2"][0])


def test_conditioning_absorbs_the_checkpoint_scale():
    """Scaling the weights in front of a gLN by s moves the chosen exponents by -log2(s) (to the nearest integer):
    the conditioned tensors -- what the fp16 pipeline stores -- have the same magnitude whatever s was."""
    from b200sep import weights
    sd, args = weights.random_state_dict(3, 4, n_blocks=2, n_repeats=1)
    base = weights.condition_state_dict(sd, args)[2]
    s = 2.0 ** 11
    sd_s = {k: v.clone() for k, v in sd.items()}
    for k in sd_s:
        if k.endswith(("shared_block.0.weight", "shared_block.0.bias", "shared_block.3.weight", "shared_block.3.bias")):
            sd_s[k] = sd_s[k] * s
    got = weights.condition_state_dict(sd_s, args)[2]
    assert [a - b for a, b in zip(base["h1"], got["h1"])] == [11, 11]
    assert [a - b for a, b in zip(base["h2"], got["h2"])] == [11, 11]
    assert base["x"] == got["x"] and base["skip"] == got["skip"] and base["enc"] == got["enc"]


def test_packed_weights_carry_the_eps_table():
    from b200sep import weights
    sd, args = weights.random_state_dict(2, 5, n_blocks=2, n_repeats=1)
    pk = weights.PackedWeights(sd, args)
    assert pk.conditioning is not None
    got = np.ctypeslib.as_array(pk.struct.gln_eps, (5,))
    assert np.array_equal(got, pk.eps) and np.all(got > 0)
    raw = weights.PackedWeights(sd, args, condition=False)
    assert raw.conditioning is None and np.allclose(np.ctypeslib.as_array(raw.struct.gln_eps, (5,)), 1e-8)
