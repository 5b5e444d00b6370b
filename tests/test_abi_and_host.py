"""CPU tests: the C-ABI library builds for sm_100a, loads, exports every symbol the header declares;
host-side logic (weights, checkpoint formats, wrapper semantics, sharding, metrics)."""
import ctypes as C
import os
import re

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_functions():
    src = open(os.path.join(ROOT, "include", "b200sep.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(b200sep_[a-z0-9_]+)\s*\(", src)))


def test_library_builds_and_exports_every_declared_symbol(built_lib):
    from b200sep import _native as nat
    declared = header_functions()
    assert len(declared) >= 12
    for name in declared:
        assert hasattr(built_lib, name), f"{name} declared in include/b200sep.h but not exported"
    assert sorted(nat.EXPORTS) == declared
    assert built_lib.b200sep_version() == 200


def test_library_is_sm100a_with_blackwell_instructions(built_lib):
    """cuobjdump is part of the image; the SASS must contain tcgen05 MMA / TMEM load / TMA load."""
    import shutil
    import subprocess
    from b200sep import _native as nat
    cu = shutil.which("cuobjdump")
    if not cu:
        pytest.skip("cuobjdump not available")
    out = subprocess.run([cu, "-sass", nat.LIB_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in out
    for mnemonic in ("UTCHMMA", "LDTM", "UTMALDG"):
        assert mnemonic in out, mnemonic


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure path")
def test_no_gpu_is_a_loud_error_not_a_fallback(built_lib):
    import b200sep
    from b200sep import _native as nat, weights
    sd, args = weights.random_state_dict(3, 0, n_blocks=1, n_repeats=1)
    packed = weights.PackedWeights(sd, args)
    cfg = packed.config("fp16")
    h = C.c_void_p()
    rc = built_lib.b200sep_model_create(C.byref(cfg), C.byref(packed.struct), 0, C.byref(h))
    assert rc != 0 and not h.value
    assert "no CUDA device" in nat.last_error() or "fallback" in nat.last_error()
    with pytest.raises(RuntimeError) as ei:
        b200sep.Separator(device="cuda", n_src=3, state_dict=sd, model_args=args)
    assert "B200" in str(ei.value)
    with pytest.raises(RuntimeError):
        b200sep.Separator(device="cpu", n_src=3, state_dict=sd, model_args=args)   # no CPU path


def test_config_validation_messages(built_lib):
    from b200sep import _native as nat, weights
    sd, args = weights.random_state_dict(3, 0, n_blocks=1, n_repeats=1)
    packed = weights.PackedWeights(sd, args)
    for field, val in (("n_src", 4), ("hid_chan", 256), ("kernel_size", 20), ("precision", 9)):
        cfg = packed.config("fp16")
        setattr(cfg, field, val)
        h = C.c_void_p()
        assert built_lib.b200sep_model_create(C.byref(cfg), C.byref(packed.struct), 0, C.byref(h)) != 0
        assert nat.last_error()
    assert built_lib.b200sep_workspace_bytes(None, None, 0) == 0
    lens = nat.i32_array([100])
    assert built_lib.b200sep_pit_workspace_bytes(lens, 1) > 0


def test_random_state_dict_matches_oracle_layout():
    from b200sep import weights
    from oracle import asteroid_restated as ar
    for n_src, kw in ((3, {}), (2, dict(kernel_size=32, stride=16, mask_act="relu"))):
        sd, args = weights.random_state_dict(n_src, 3, **kw)
        m = ar.ConvTasNet(**args)
        ref = m.state_dict()
        assert set(sd) == set(ref)
        assert all(tuple(sd[k].shape) == tuple(ref[k].shape) for k in ref)
        m.load_state_dict(sd)           # strict
        assert weights.model_args_from_state_dict(sd)["n_src"] == n_src
        assert weights.model_args_from_state_dict(sd)["kernel_size"] == args["kernel_size"]


def test_checkpoint_formats(tmp_path):
    from b200sep import weights
    sd, args = weights.random_state_dict(2, 1, n_blocks=2, n_repeats=1, mask_act="relu")
    p1, p2, p3 = (str(tmp_path / n) for n in ("asteroid.bin", "wrapped.pt", "bare.pt"))
    torch.save({"model_name": "ConvTasNet", "state_dict": sd, "model_args": args}, p1)
    torch.save({"state_dict": sd}, p2)
    torch.save(sd, p3)
    a1 = weights.load_checkpoint(p1)[1]
    assert a1["mask_act"] == "relu" and a1["n_blocks"] == 2
    for p in (p2, p3):                       # no model_args: hyper-parameters derived from shapes
        s, a = weights.load_checkpoint(p)
        assert a["n_src"] == 2 and a["n_blocks"] * a["n_repeats"] == 2 and a["mask_act"] == "sigmoid"
        assert torch.equal(s["masker.mask_net.1.bias"], sd["masker.mask_net.1.bias"])
    with pytest.raises(FileNotFoundError):
        weights.load_checkpoint(str(tmp_path / "nope.pt"))
    bad = dict(sd)
    del bad["masker.TCN.1.res_conv.bias"]
    with pytest.raises(RuntimeError):
        weights.PackedWeights(bad, args)


def test_packed_weights_point_at_contiguous_fp32_copies():
    from b200sep import weights
    sd, args = weights.random_state_dict(3, 2, n_blocks=2, n_repeats=2)
    sd["masker.bottleneck.1.weight"] = sd["masker.bottleneck.1.weight"].double()   # any float dtype accepted
    pk = weights.PackedWeights(sd, args, condition=False)
    w = pk.struct
    assert not w.gln_eps or np.allclose(np.ctypeslib.as_array(w.gln_eps, (9,)), 1e-8)
    assert np.ctypeslib.as_array(w.enc_filters, (512 * 16,))[5] == pytest.approx(float(sd["encoder.filterbank._filters"].reshape(-1)[5]))
    assert np.ctypeslib.as_array(w.res_bias[3], (128,))[7] == pytest.approx(float(sd["masker.TCN.3.res_conv.bias"][7]))
    cfg = pk.config("fp32")
    assert (cfg.n_blocks, cfg.n_repeats, cfg.precision, cfg.mask_act) == (2, 2, 0, 0)
    with pytest.raises(RuntimeError):
        pk.config("int8")


def test_ensure_sr_matches_reference_semantics():
    """separation.py:91-103 -- linear interpolation in float64, output float32, rounding of the length."""
    from b200sep.separator import Separator
    s = Separator.__new__(Separator)
    s.sample_rate = 16000
    x = np.random.default_rng(0).standard_normal(8000).astype(np.float32)
    assert s._ensure_sr(x, 16000) is x
    y = s._ensure_sr(x, 8000)
    assert y.shape == (16000,) and y.dtype == np.float32      # the SURVEY a4 "trap"
    assert y[0] == x[0] and y[-1] == x[-1]
    assert s._ensure_sr(x[:1], 8000).shape == (1,)
    from oracle import shim
    if shim.reference_available():
        Ref = shim.import_reference_separator()
        r = Ref.__new__(Ref)
        r.sample_rate = 16000
        assert np.array_equal(r._ensure_sr(x, 8000), y)
        assert np.array_equal(r._ensure_sr(x, 44100), s._ensure_sr(x, 44100))


def test_lpt_partition_and_buckets():
    from b200sep import sharding, synth
    lens = synth.segment_lengths(10000, 5)
    costs = [sharding.frames_of(n) for n in lens]
    for world in (1, 2, 4, 8):
        shards = sharding.lpt_partition(costs, world)
        flat = sorted(i for s in shards for i in s)
        assert flat == list(range(len(lens)))                       # a partition
        loads = [sum(costs[i] for i in s) for s in shards]
        assert max(loads) - min(loads) <= max(costs)                # LPT balance bound
        assert shards == sharding.lpt_partition(costs, world)       # deterministic
        assert shards == sharding._lpt_partition_py(costs, world)   # the native helper follows the stated rule exactly
    b = sharding.bucket_by_length(range(200), lens, 400000)
    assert sorted(i for x in b for i in x) == list(range(200))
    assert all(sum(lens[i] for i in x) <= 400000 or len(x) == 1 for x in b)
    flat_lens = [lens[i] for x in b for i in x]
    assert flat_lens == sorted(flat_lens)
    assert sharding.frames_of(32000) == 3999 and sharding.frames_of(1600) == 199
    with pytest.raises(ValueError):
        sharding.lpt_partition([1], 0)
    assert sharding.lpt_partition([], 3) == [[], [], []]
    assert sharding.lpt_partition([5, 5, 5, 5], 2) == sharding._lpt_partition_py([5, 5, 5, 5], 2) == [[0, 2], [1, 3]]   # ties


def test_pit_and_partition_argument_validation(built_lib):
    """Error paths that return before any CUDA call (so they run without a GPU)."""
    from b200sep import _native as nat
    lens = nat.i32_array([1000, 2000])
    assert built_lib.b200sep_pit_workspace_bytes_n(lens, 2, 9) == 0 and "at most 8" in nat.last_error()
    assert built_lib.b200sep_pit_workspace_bytes_n(lens, 2, 8) > built_lib.b200sep_pit_workspace_bytes_n(lens, 2, 3) \
        == built_lib.b200sep_pit_workspace_bytes(lens, 2)          # one block of partial moments per group of three
    dummy = (C.c_double * 64)()
    p = C.cast(dummy, C.c_void_p)
    for K, N in ((1, 3), (4, 4), (3, 2), (2, 9)):
        rc = built_lib.b200sep_pit_sisdr_batch(p, p, p, None, lens, 2, K, N, 0, p, p, p, p, None, p, C.c_size_t(1 << 20), None)
        assert rc != 0 and "K in {2,3}" in nat.last_error(), (K, N, nat.last_error())
    owner = (C.c_int32 * 4)()
    costs = (C.c_int64 * 4)(5, 1, 5, 3)
    assert built_lib.b200sep_lpt_partition(C.cast(costs, C.c_void_p), 4, 0, C.cast(owner, C.c_void_p)) != 0
    assert built_lib.b200sep_lpt_partition(C.cast(costs, C.c_void_p), 4, 2, C.cast(owner, C.c_void_p)) == 0
    assert list(owner) == [0, 1, 1, 0]    # 5 -> r0, 5 -> r1, 3 -> r0 (tie: lowest rank), 1 -> r1


def test_metrics_match_reference_aggregation():
    from b200sep import metrics
    vals = [10.1234567, 12.5, float("nan"), 9.75]
    imp = [13.0, 14.25, 1.0, float("nan")]
    m = metrics.sep_metrics(vals, imp)
    assert m["sep_eval_segments"] == 2 and m["sep_eval_k_refs"] is None
    arr = np.asarray([10.1234567, 12.5], np.float32)
    assert m["sep_sisdr_mean"] == round(float(np.mean(arr)), 4)
    assert m["sep_sisdr_std"] == round(float(np.std(arr)), 4)          # population std
    assert m["sep_sisdri_median"] == round(float(np.median(np.float32([13.0, 14.25]))), 4)
    assert metrics.agg_stats([]) == {"mean": None, "median": None, "std": None, "count": 0}
    row = metrics.sep_detail_row("a/mix.wav", 1.23456, 2.5, 3, 10.16, 13.54321, [1, 2, 0])
    assert row == ["a/mix.wav", "1.235", "2.500", 3, "10.1600", "13.5432", "1;2;0"]
    from oracle import pit_restated as pr
    assert pr.aggregate_stats([1.0, 2.0, 4.0])[2] == pytest.approx(float(np.std(np.float32([1, 2, 4]))))


def test_synthetic_workload_is_deterministic_and_shaped():
    from b200sep import synth
    lens = synth.segment_lengths(64, 2)
    assert len(lens) == 64 and min(lens) >= 1600 and max(lens) <= 32000
    assert 120 * 8000 < sum(lens) < 150 * 8000                       # ~134 audio-seconds
    m1, s1 = synth.make_mixture(3, 4000, 3)
    m2, s2 = synth.make_mixture(3, 4000, 3)
    assert np.array_equal(m1, m2) and m1.dtype == np.float32 and len(s1) == 3
    assert np.allclose(np.sum(s1, 0), m1, atol=1e-6) and np.abs(m1).max() <= 0.9 + 1e-6
    assert synth.streaming_chunk_samples(2.0) == 15360 and synth.streaming_chunk_samples(5.0) == 39936
