"""GPU parity of SURVEY.md 8(f) rows 1-3 (through the C ABI): device resampling, frame RMS / GT overlap masks and the
batched pipeline driver, against the committed golden vectors (reference functions / torchaudio) and the oracle.

Tolerances: linear resample bit-exact (fp64 replay of np.interp); sinc resample <= 2e-6 absolute (fp32 FIR, different
summation order than torch's conv1d); frame RMS <= 2e-6 relative (fp64 accumulation vs NumPy's fp32 pairwise sum);
GT overlap mask identical except frames whose RMS lies within that tolerance of the threshold."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import aux_restated as ax  # noqa: E402
from oracle import asteroid_restated as ar  # noqa: E402
from oracle import pit_restated as pr  # noqa: E402


@pytest.fixture(scope="module")
def b200():
    import b200sep
    if not torch.cuda.is_available():
        pytest.fail("GPU tests need a B200")
    return b200sep


def _cases(g, prefix):
    i = 0
    while f"{prefix}{i}_cfg" in g:
        yield i
        i += 1


def test_linear_resample_bit_exact_against_reference_golden(b200, golden_dir):
    from b200sep import aux
    g = np.load(os.path.join(golden_dir, "aux_golden.npz"))
    for i in _cases(g, "lin"):
        sr, tgt = (int(v) for v in g[f"lin{i}_cfg"])
        y = aux.resample_batch([g[f"lin{i}_in"]], sr, tgt, "linear")[0]
        assert y.shape == g[f"lin{i}_out"].shape and np.array_equal(y, g[f"lin{i}_out"]), i
    # ragged batch, many rates, against the oracle (= the reference's np.interp call) -- bit-exact
    rng = np.random.default_rng(7)
    for sr_in, sr_out in [(16000, 8000), (8000, 16000), (44100, 8000), (11025, 8000), (48000, 16000)]:
        sigs = [rng.standard_normal(int(n)).astype(np.float32) for n in rng.integers(1, 40000, 24)] + [np.zeros(1, np.float32)]
        out = aux.resample_batch(sigs, sr_in, sr_out, "linear")
        for x, y in zip(sigs, out):
            r = ax.ensure_sr_linear(x, sr_in, sr_out)
            assert y.shape == r.shape and np.array_equal(y, r)


def test_sinc_resample_against_torchaudio_golden(b200, golden_dir):
    from b200sep import aux
    g = np.load(os.path.join(golden_dir, "aux_golden.npz"))
    for i in _cases(g, "sinc"):
        sr, tgt = (int(v) for v in g[f"sinc{i}_cfg"])
        y = aux.resample_batch([g[f"sinc{i}_in"]], sr, tgt, "sinc")[0]
        r = g[f"sinc{i}_out"]
        assert y.shape == r.shape
        assert float(np.abs(y - r).max()) <= 2e-6, (i, float(np.abs(y - r).max()))
    rng = np.random.default_rng(8)
    sigs = [rng.standard_normal(int(n)).astype(np.float32) * 0.1 for n in rng.integers(2, 30000, 16)]
    for sr_in, sr_out in [(16000, 8000), (44100, 16000), (8000, 16000)]:
        out = aux.resample_batch(sigs, sr_in, sr_out, "sinc")
        for x, y in zip(sigs, out):
            r = ax.resample_sinc(x, sr_in, sr_out)
            assert y.shape == r.shape and float(np.abs(y - r).max()) <= 2e-6


def test_frame_rms_and_gt_overlap_mask_against_reference_golden(b200, golden_dir):
    from b200sep import aux
    g = np.load(os.path.join(golden_dir, "aux_golden.npz"))
    pairs, want = [], []
    for i in _cases(g, "rms"):
        sr, win, hop, thr = g[f"rms{i}_cfg"]
        s1, s2 = g[f"rms{i}_s1"], g[f"rms{i}_s2"]
        r1 = aux.frame_rms(s1, int(sr), win, hop)
        assert r1.shape == g[f"rms{i}_rms1"].shape
        np.testing.assert_allclose(r1, g[f"rms{i}_rms1"], rtol=2e-6, atol=0)
        m = aux.build_gt_overlap_mask(s1, s2, int(sr), win, hop, thr)
        ref_m = g[f"rms{i}_mask"]
        assert m.shape == ref_m.shape and m.dtype == bool
        if not np.array_equal(m, ref_m):  # only threshold-tie frames may differ
            r2 = g[f"rms{i}_rms2"]
            lim = max(float(g[f"rms{i}_rms1"].max()), float(r2.max()), 1e-9) * thr
            d = np.flatnonzero(m != ref_m)
            assert all(min(abs(g[f"rms{i}_rms1"][k] - lim), abs(r2[k] - lim)) <= 4e-6 * lim for k in d)
    # batched form over a ragged set of pairs == one call per pair; silent and short signals
    rng = np.random.default_rng(9)
    pairs = [(rng.standard_normal(n).astype(np.float32), rng.standard_normal(n).astype(np.float32) * 0.3)
             for n in (16000, 50, 200, 4321, 199)] + [(np.zeros(800, np.float32), np.zeros(800, np.float32))]
    got = aux.build_gt_overlap_mask_batch(pairs, 8000, 0.025, 0.010, 0.05)
    for (a, b), m in zip(pairs, got):
        assert np.array_equal(m, ax.build_gt_overlap_mask(a, b, 8000, 0.025, 0.010, 0.05))
    assert got[-1].all()  # silent pair: rms = sqrt(1e-12) on every frame > 5 % of itself (the reference's own behaviour)
    with pytest.raises(ValueError, match="win too small"):
        aux.frame_rms(pairs[0][0], 8000, 0.00001, 0.01)


def test_separator_resamples_on_the_device_like_ensure_sr(b200):
    from b200sep import synth, weights
    sd, args = weights.random_state_dict(3, 0)
    sep = b200.Separator(backend="b200", device="cuda:0", sample_rate=8000, n_src=3, state_dict=sd, model_args=args)
    mixes = [synth.make_mixture(70 + i, n, 3, sr=16000)[0] for i, n in enumerate((16000, 9001, 3300))]
    dev = sep.separate_batch(mixes, 16000)                                   # device _ensure_sr
    host = sep.engine.separate_batch([sep._ensure_sr(m, 16000) for m in mixes])  # NumPy _ensure_sr, then the same forward
    for m, a, b in zip(mixes, dev, host):
        n = int(round(len(m) * 8000 / 16000))
        assert all(x.shape == (n,) for x in a)
        assert all(np.array_equal(x, y) for x, y in zip(a, b))  # identical resampled input -> identical forward
    one = sep.separate(mixes[1], 16000)
    assert all(np.array_equal(x, y) for x, y in zip(one, host[1]))


def test_batched_driver_end_to_end_matches_per_segment_reference_flow(b200):
    """The driver (one separate_batch + one pit_sisdr_batch) against the reference's flow: one separate() per overlap
    segment by the ORACLE model, scored by the oracle's restatement of the reference PIT."""
    from b200sep import synth, weights
    from b200sep.pipeline import MixtureJob, OverlapBatchDriver
    sd, args = weights.random_state_dict(3, 0)
    sep = b200.Separator(backend="b200", device="cuda:0", sample_rate=8000, n_src=3, state_dict=sd, model_args=args)
    ref_model = ar.ConvTasNet(**args).eval()
    ref_model.load_state_dict(sd)
    jobs = []
    for j in range(2):
        mix, srcs = synth.make_mixture(90 + j, 8000 * 5, 3)
        segs = [(0.0, 0.6, False), (0.6, 1.9, True), (1.9, 2.1, True), (2.1, 4.35, True), (4.35, 5.0, False)]
        jobs.append(MixtureJob(mix=mix, sr=8000, segments=segs, mix_path=f"m{j}.wav", refs=srcs))
    res = OverlapBatchDriver(sep, min_overlap_dur=0.4, eval_separation=True).run(jobs)
    assert len(res.overlap) == 4 and len(res.clean) == 6 and len(res.sep_details_rows) == 4
    k = 0
    for job in jobs:
        for s, e, ol in job.segments:
            if not ol or e - s < 0.4:
                continue
            s_i, e_i = int(s * 8000), int(e * 8000)
            chunk = job.mix[s_i:e_i]
            with torch.no_grad():
                ref = ref_model(torch.from_numpy(chunk)[None])[0].numpy()
            rec = res.overlap[k]
            assert min(pr.si_sdr_f64(ref[i], rec.branches[i]) for i in range(3)) >= 40.0
            best, sdri, idx = pr.compute_sdr_improvement_pit_k_o3(chunk, [r[s_i:e_i] for r in job.refs], list(ref))
            assert rec.perm == idx and abs(rec.sisdri - sdri) <= 0.05 and abs(rec.sisdr - best) <= 0.05
            assert res.sep_details_rows[k][:4] == [job.mix_path, f"{s:.3f}", f"{e:.3f}", 3]
            k += 1
    m = res.sep_metrics()
    assert m["sep_eval_segments"] == 4 and m["sep_sisdri_mean"] is not None
    # the fused (device-resident) evaluation gives the same rows
    res_f = OverlapBatchDriver(sep, min_overlap_dur=0.4, eval_separation=True, fused_eval=True).run(jobs)
    assert [r[:4] + [r[6]] for r in res_f.sep_details_rows] == [r[:4] + [r[6]] for r in res.sep_details_rows]
    assert np.allclose(res_f.sep_sisdri_list, res.sep_sisdri_list, atol=1e-6)
    for a, b in zip(res_f.overlap, res.overlap):
        assert all(np.array_equal(x, y) for x, y in zip(a.branches, b.branches))


def test_separate_and_score_on_device_equals_the_two_step_path(b200):
    """--eval-separation with estimates and references resident on the device == separate_batch + pit_sisdr_batch."""
    import time
    from b200sep import synth, weights
    sd, args = weights.random_state_dict(3, 0)
    sep = b200.Separator(backend="b200", device="cuda:0", sample_rate=8000, n_src=3, state_dict=sd, model_args=args)
    mixes, srcs = synth.make_batch(40, seed=11, n_src=3, lo_s=0.4, hi_s=4.0, with_sources=True)
    outs = sep.separate_batch(mixes, 8000)
    two = b200.pit_sisdr_batch(mixes, srcs, outs, "overlap3")
    one, wav = sep.engine.separate_and_score(mixes, srcs, "overlap3")
    assert np.array_equal(one.perm, two.perm)
    np.testing.assert_allclose(one.sisdr, two.sisdr, rtol=0, atol=1e-9)
    np.testing.assert_allclose(one.sisdri, two.sisdri, rtol=0, atol=1e-9)
    for a, b in zip(wav, outs):
        assert all(np.array_equal(x, y) for x, y in zip(a, b))
    # K = 2 references against 3 estimates, records only
    r2, none = sep.engine.separate_and_score(mixes[:7], [s[:2] for s in srcs[:7]], "overlap3", return_waveforms=False)
    t2 = b200.pit_sisdr_batch(mixes[:7], [s[:2] for s in srcs[:7]], outs[:7], "overlap3")
    assert none is None and np.array_equal(r2.perm, t2.perm)
    np.testing.assert_allclose(r2.sisdri, t2.sisdri, rtol=0, atol=1e-9)
    with pytest.raises(ValueError):
        sep.engine.separate_and_score(mixes[:2], [[s[0][:-1], s[1], s[2]] for s in srcs[:2]])
    t0 = time.perf_counter(); sep.engine.separate_and_score(mixes, srcs, "overlap3", return_waveforms=False); t1 = time.perf_counter()
    outs = sep.separate_batch(mixes, 8000); b200.pit_sisdr_batch(mixes, srcs, outs, "overlap3"); t2_ = time.perf_counter()
    print(f"separate_and_score {1e3 * (t1 - t0):.2f} ms vs two-step {1e3 * (t2_ - t1):.2f} ms for 40 segments")


def test_stream_apis_match_the_one_at_a_time_calls(b200):
    """separate_stream / score_stream keep two batches in flight (own stream, scratch and pinned staging each): results
    must be bit-identical to the synchronous calls and come back in order, whatever the batch sizes."""
    from b200sep import synth, weights
    sd, args = weights.random_state_dict(3, 0)
    sep = b200.Separator(backend="b200", device="cuda:0", sample_rate=8000, n_src=3, state_dict=sd, model_args=args)
    batches, refs = [], []
    for bi, n_seg in enumerate((5, 1, 9, 3, 7)):
        lens = synth.segment_lengths(n_seg, 70 + bi, 0.2, 2.0)
        ms = [synth.make_mixture(800 + 10 * bi + i, n, 3) for i, n in enumerate(lens)]
        batches.append([m for m, _ in ms])
        refs.append([s for _, s in ms])
    want = [sep.separate_batch(b, 8000) for b in batches]
    got = list(sep.separate_stream(batches, 8000))
    assert len(got) == len(want)
    for w, g in zip(want, got):
        assert len(w) == len(g)
        for ws, gs in zip(w, g):
            for a, b_ in zip(ws, gs):
                assert np.array_equal(a, b_)
    one = [sep.engine.separate_and_score(b, r, "overlap3", return_waveforms=False)[0] for b, r in zip(batches, refs)]
    two = [r for r, _ in sep.engine.score_stream(zip(batches, refs), "overlap3")]
    for a, b_ in zip(one, two):
        assert np.array_equal(a.sisdr, b_.sisdr) and np.array_equal(a.sisdri, b_.sisdri)
        assert np.array_equal(a.perm, b_.perm) and np.array_equal(a.gap, b_.gap)
    sep.engine.close()
