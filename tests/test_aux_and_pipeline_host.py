"""CPU tests of SURVEY.md 8(f): the aux oracle against the committed golden vectors (made from the reference's own
functions and from torchaudio), the host-only length rules of the C ABI, the batched pipeline driver's selection /
row rules against a per-segment loop written like the reference's, and the result writers."""
import csv
import json
import os

import numpy as np
import pytest

from oracle import aux_restated as ax
from oracle import pit_restated as pr


def _cases(g, prefix):
    i = 0
    while f"{prefix}{i}_cfg" in g:
        yield i
        i += 1


def test_aux_oracle_matches_golden(golden_dir):
    g = np.load(os.path.join(golden_dir, "aux_golden.npz"))
    n = 0
    for i in _cases(g, "lin"):
        sr, tgt = (int(v) for v in g[f"lin{i}_cfg"])
        y = ax.ensure_sr_linear(g[f"lin{i}_in"], sr, tgt)
        assert np.array_equal(y, g[f"lin{i}_out"])  # bit-exact (fp64 interpolation rounded once)
        n += 1
    for i in _cases(g, "sinc"):
        sr, tgt = (int(v) for v in g[f"sinc{i}_cfg"])
        y = ax.resample_sinc(g[f"sinc{i}_in"], sr, tgt)
        r = g[f"sinc{i}_out"]
        assert y.shape == r.shape
        assert float(np.abs(y - r).max()) <= 1e-6  # fp32 filter bank, different summation order
        n += 1
    for i in _cases(g, "rms"):
        sr, win, hop, thr = g[f"rms{i}_cfg"]
        s1, s2 = g[f"rms{i}_s1"], g[f"rms{i}_s2"]
        assert np.array_equal(ax.frame_rms(s1, int(sr), win, hop), g[f"rms{i}_rms1"])
        assert np.array_equal(ax.frame_rms(s2, int(sr), win, hop), g[f"rms{i}_rms2"])
        assert np.array_equal(ax.build_gt_overlap_mask(s1, s2, int(sr), win, hop, thr), g[f"rms{i}_mask"])
        n += 1
    assert n >= 15


def test_length_rules_of_the_c_abi_match_the_reference_formulas(built_lib):
    import math
    from b200sep import aux
    rng = np.random.default_rng(0)
    lens = [0, 1, 2, 3, 4, 5, 7, 8, 15, 16, 17, 1599, 1600, 4001, 32000] + [int(v) for v in rng.integers(1, 200000, 200)]
    for sr_in, sr_out in [(16000, 8000), (8000, 16000), (44100, 8000), (22050, 16000), (48000, 16000), (8000, 8000), (11025, 8000)]:
        got = aux.resample_out_lens(lens, sr_in, sr_out, "linear")
        for n, o in zip(lens, got):
            x = np.zeros(n, np.float32)
            assert o == ax.ensure_sr_linear(x, sr_in, sr_out).shape[0], (n, sr_in, sr_out)
        got = aux.resample_out_lens(lens, sr_in, sr_out, "sinc")
        g = math.gcd(sr_in, sr_out)
        for n, o in zip(lens, got):
            want = n if (sr_in == sr_out or n <= 1) else int(math.ceil((sr_out // g) * n / (sr_in // g)))
            assert o == want, (n, sr_in, sr_out)
    for win_s, hop_s in [(200, 80), (400, 160), (1, 1), (800, 400)]:
        got = aux.frame_rms_out_lens(lens, win_s, hop_s)
        for n, o in zip(lens, got):
            assert o == len(range(0, max(n - win_s + 1, 1), hop_s))
    with pytest.raises(ValueError, match="win too small"):
        aux.frame_rms_out_lens([10], 0, 1)
    with pytest.raises(ValueError, match="hop too small"):
        aux.frame_rms_out_lens([10], 4, 0)


class _FakeSep:
    """Deterministic stand-in with the Separator.separate / separate_batch contract (no GPU)."""
    n_src = 3

    def __init__(self):
        self.batch_calls = 0

    def separate(self, chunk, sr):
        c = np.asarray(chunk, np.float32)
        return [c * 0.5, np.roll(c, 3) * 0.25, (c ** 2 - np.mean(c ** 2)).astype(np.float32)]

    def separate_batch(self, chunks, sr):
        self.batch_calls += 1
        return [self.separate(c, sr) for c in chunks]


def _oracle_pit_batch(mixes, refs, preds, variant="overlap3"):
    from b200sep.pit import PitResult
    n, K = len(mixes), len(refs[0])
    sisdr, sisdri, perm = np.full(n, np.nan), np.full(n, np.nan), np.full((n, K), -1, np.int32)
    for i in range(n):
        b, d, idx = pr.compute_sdr_improvement_pit_k_o3(mixes[i], list(refs[i]), list(preds[i]))
        if idx:
            sisdr[i], sisdri[i], perm[i] = b, d, idx
    return PitResult(sisdr, sisdri, perm, np.zeros(n, bool), np.zeros((n, 3, 4)))


def _reference_style_loop(sep, jobs, min_overlap_dur):
    """The per-segment loop of overlap3_core.py:604-744 reduced to the separation + --eval-separation parts."""
    from b200sep import metrics as M
    rows, sisdr, sisdri, branches_all, n_clean = [], [], [], [], 0
    for job in jobs:
        sr, mix_np = job.sr, job.mix
        for s, e, is_olap in job.segments:
            if e - s <= 0:
                continue
            s_i, e_i = int(s * sr), int(e * sr)
            chunk = mix_np[s_i:e_i]
            if (not is_olap) or (e - s) < min_overlap_dur:
                n_clean += 1
                continue
            branches = list(sep.separate(chunk, sr))
            branches_all.append(branches)
            if job.refs:
                K = 3 if len(job.refs) >= 3 else len(job.refs)
                if K not in (2, 3) or len(branches) < K:
                    continue
                refs = [np.asarray(r)[s_i:e_i] for r in job.refs[:K]]
                best, sdri, idx = pr.compute_sdr_improvement_pit_k_o3(chunk, refs, [np.asarray(b, np.float32) for b in branches])
                if not (np.isnan(best) or np.isnan(sdri)):
                    sisdr.append(float(best)); sisdri.append(float(sdri))
                    rows.append(M.sep_detail_row(job.mix_path, s, e, K, best, sdri, idx))
    return rows, sisdr, sisdri, branches_all, n_clean


def test_batched_driver_reproduces_the_per_segment_loop():
    from b200sep import synth
    from b200sep.pipeline import MixtureJob, OverlapBatchDriver
    jobs = []
    for j in range(3):
        mix, srcs = synth.make_mixture(50 + j, 8000 * 6, 3)
        segs = [(0.0, 0.5, False), (0.5, 0.8, True), (0.8, 0.8, True), (1.0, 2.25, True), (2.25, 2.0, True),
                (2.5, 2.9001, True), (3.0, 5.5, True), (5.5, 6.0, False), (5.9, 6.5, True)]
        refs = srcs if j != 1 else srcs[:2]  # one mixture with only two references (K = 2)
        jobs.append(MixtureJob(mix=mix, sr=8000, segments=segs, mix_path=f"mix{j}.wav", refs=refs))
    jobs.append(MixtureJob(mix=jobs[0].mix, sr=8000, segments=[(0.2, 1.2, True)], mix_path="noref.wav", refs=None))
    sep = _FakeSep()
    drv = OverlapBatchDriver(sep, min_overlap_dur=0.4, eval_separation=True, pit=_oracle_pit_batch)
    res = drv.run(jobs)
    rows, sisdr, sisdri, branches_all, n_clean = _reference_style_loop(_FakeSep(), jobs, 0.4)
    assert sep.batch_calls == 1  # every overlap chunk of every mixture in ONE call
    assert len(res.overlap) == len(branches_all) and len(res.clean) == n_clean
    for r, b in zip(res.overlap, branches_all):
        assert all(np.array_equal(x, y) for x, y in zip(r.branches, b))
    assert res.sep_details_rows == rows
    assert res.sep_sisdr_list == sisdr and res.sep_sisdri_list == sisdri
    assert res.counters["n_seen_overlap_segments"] == len(branches_all)
    assert abs(res.counters["total_overlap_audio_sec"] - sum(r.e - r.s for r in res.overlap)) < 1e-12
    m = res.sep_metrics()
    assert m["sep_eval_segments"] == len(rows) and m["sep_sisdr_mean"] == round(float(np.mean(np.asarray(sisdr, np.float32))), 4)
    # a batch-size cap splits the work without changing results
    sep2 = _FakeSep()
    res2 = OverlapBatchDriver(sep2, 0.4, True, pit=_oracle_pit_batch, max_batch_samples=30000).run(jobs)
    assert sep2.batch_calls > 1 and res2.sep_details_rows == rows


def test_result_writers_match_the_reference_formats(tmp_path):
    from b200sep import metrics as M
    rows = [M.sep_detail_row("a,b.wav", 0.5, 2.25, 3, 10.12345, 13.5, [1, 2, 0]), M.sep_detail_row("c.wav", 1, 2, 2, -0.00004, 0.1, [0, 1])]
    p = tmp_path / "overlap_sep_details.csv"
    M.write_sep_details_csv(str(p), rows)
    with open(p, newline="", encoding="utf-8") as fh:
        got = list(csv.reader(fh))
    assert got[0] == ["wav", "start", "end", "k_refs", "sisdr", "sisdri", "selected_pred_indices"]
    assert got[1] == ["a,b.wav", "0.500", "2.250", "3", "10.1235", "13.5000", "1;2;0"]
    assert got[2] == ["c.wav", "1.000", "2.000", "2", "-0.0000", "0.1000", "0;1"]
    met = M.merge_metrics({"total_audio_sec": 12.0}, [10.0, 12.0, float("nan")], [13.0, 15.0, 1.0], time_sep_sec=0.12345)
    assert met["time_sep_sec"] == 0.123 and met["sep_eval_segments"] == 2 and met["sep_sisdr_mean"] == 11.0
    assert met["sep_sisdri_std"] == 1.0 and met["sep_eval_k_refs"] is None
    q = tmp_path / "metrics.json"
    M.write_metrics_json(str(q), met)
    assert json.load(open(q, encoding="utf-8")) == met
