"""Generates the committed golden fixtures under tests/golden/ (run in the BUILD CONTAINER, where
/root/reference exists; the fixtures travel, the reference does not).

  pit_golden.npz        -- outputs of the REFERENCE's own PIT functions (imported unmodified from
                           /root/reference/scripts/osd/{overlap3_core,evaluate_with_sources}.py)
                           on seeded inputs: pins oracle/pit_restated.py and the CUDA PIT kernel.
  pit_golden_many.npz   -- the same functions with more estimates than references (N = 4 .. 8, K = 2 / 3), same keys as
                           pit_golden_ties.npz.
  pit_golden_ties.npz   -- the same functions on exact ties and near-ties (duplicated estimates / references, all
                           scores equal, 1e-7-relative differences) + the top-2 gap of the reference's pair matrix.
  separator_golden.npz  -- output of the REFERENCE's own Separator wrapper
                           (/root/reference/src/osd/separation.py, unmodified) driving the restated
                           Asteroid model through the oracle shim on seeded mixtures and the seeded
                           random-init weights of b200sep.weights.random_state_dict.  The model
                           arithmetic itself stays "parity unpinned" (asteroid is not on the box).

    python -m oracle.make_golden
"""
from __future__ import annotations

import hashlib
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import asteroid_restated as ar  # noqa: E402
from oracle import shim  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def state_dict_digest(sd) -> str:
    h = hashlib.sha256()
    for k in sorted(sd):
        h.update(k.encode())
        h.update(np.ascontiguousarray(sd[k].detach().cpu().numpy(), dtype=np.float32).tobytes())
    return h.hexdigest()


def pit_cases():
    """Seeded (mix, refs, preds) cases: clean permutations, noisy, K=2/N=3, ragged lengths,
    near-silent and all-zero signals, a short (200-sample) segment."""
    rng = np.random.default_rng(20240917)
    cases = []
    for ci in range(18):
        K = 3 if ci % 3 != 1 else 2
        N = 3 if (K == 3 or ci % 2 == 0) else 2
        n = int(rng.integers(200, 3000))
        refs = [(rng.standard_normal(n) * 10 ** rng.uniform(-2.5, -1)).astype(np.float32) for _ in range(K)]
        extra = [(rng.standard_normal(n) * 0.03).astype(np.float32) for _ in range(N - K)]
        order = rng.permutation(N)
        noise = 10 ** rng.uniform(-2.5, -0.3)
        base = refs + extra
        preds = [(base[j] + noise * rng.standard_normal(n).astype(np.float32) * float(np.std(base[j]) + 1e-3)).astype(np.float32)
                 for j in order]
        mix = np.sum(refs, axis=0).astype(np.float32)
        if ci == 5:
            preds[0] = np.zeros(n, np.float32)           # a silent estimate
        if ci == 7:
            refs[1] = np.zeros(n, np.float32)            # a silent reference
        if ci == 9:
            preds = [p[: n - 37] for p in preds]          # estimates shorter than references
        if ci == 11:
            mix = (mix + 0.25).astype(np.float32)         # DC offset
        cases.append((K, N, mix, refs, preds))
    return cases


def make_pit_golden():
    o3, ews = shim.import_reference_pit()
    out = {}
    cases = pit_cases()
    out["n_cases"] = np.int64(len(cases))
    for ci, (K, N, mix, refs, preds) in enumerate(cases):
        out[f"c{ci}_KN"] = np.array([K, N], np.int64)
        out[f"c{ci}_mix"] = mix
        for k, r in enumerate(refs):
            out[f"c{ci}_ref{k}"] = r
        for j, p in enumerate(preds):
            out[f"c{ci}_pred{j}"] = p
        # ---- overlap3_core.py
        pair = np.array([[o3._si_sdr(r, p) for p in preds] for r in refs], np.float64)
        base = np.array([o3._si_sdr(r, mix) for r in refs], np.float64)
        best, sdri, idx = o3._compute_sdr_improvement_pit_k(mix, refs, preds)
        out[f"c{ci}_o3_pair"] = pair
        out[f"c{ci}_o3_base"] = base
        out[f"c{ci}_o3_res"] = np.array([best, sdri], np.float64)
        out[f"c{ci}_o3_idx"] = np.array(idx, np.int64)
        # ---- evaluate_with_sources.py (operates on equal-length arrays at its call site)
        m = min([len(mix)] + [len(r) for r in refs] + [len(p) for p in preds])
        mixc, refsc, predsc = mix[:m], [r[:m] for r in refs], [p[:m] for p in preds]
        pair = np.array([[ews.si_sdr(r, p) for p in predsc] for r in refsc], np.float64)
        out[f"c{ci}_ews_pair"] = pair
        if K == 2:
            best, sdri, idx, sw = ews.compute_sdr_improvement_pit(mixc, refsc[0], refsc[1], predsc)
            out[f"c{ci}_ews_swapped"] = np.int64(bool(sw))
        else:
            best, sdri, idx = ews.compute_sdr_improvement_pit_k(mixc, refsc, predsc)
        out[f"c{ci}_ews_res"] = np.array([best, sdri], np.float64)
        out[f"c{ci}_ews_idx"] = np.array(idx, np.int64)
    np.savez_compressed(os.path.join(OUT, "pit_golden.npz"), **out)
    print("wrote pit_golden.npz with", len(cases), "cases")


def pit_tie_cases():
    """Exact ties and near-ties: duplicated estimates, duplicated references, all-equal scores,
    estimates / references that differ by 1e-7 .. 1e-6 relative."""
    rng = np.random.default_rng(20261018)
    n = 2400

    def sig(scale=0.05):
        return (rng.standard_normal(n) * scale).astype(np.float32)

    def noisy(x, a):
        return (x + a * float(np.std(x)) * rng.standard_normal(n).astype(np.float32)).astype(np.float32)

    cases = []
    r = [sig(), sig(), sig()]
    e0 = noisy(r[0], 0.2)
    cases.append((3, 3, [r[0], r[1], r[2]], [e0, e0.copy(), noisy(r[2], 0.2)]))                  # t0 duplicated estimates
    cases.append((3, 3, [r[0], r[0].copy(), r[2]], [noisy(r[0], 0.3), noisy(r[0], 0.1), noisy(r[2], 0.2)]))  # t1 dup refs
    cases.append((3, 3, [r[0], r[1], r[2]], [e0, e0.copy(), e0.copy()]))                          # t2 all estimates equal
    cases.append((3, 3, [r[1], r[1].copy(), r[1].copy()], [e0, e0.copy(), e0.copy()]))            # t3 every score equal
    cases.append((3, 3, [r[0], r[1], r[2]], [e0, (e0 + np.float32(1e-7) * sig(1.0)).astype(np.float32), noisy(r[2], 0.2)]))  # t4 near tie
    cases.append((2, 2, [r[0], r[1]], [e0, e0.copy()]))                                            # t5 K=2 duplicated estimates
    cases.append((2, 3, [r[0], r[1]], [noisy(r[1], 0.2), e0, e0.copy()]))                          # t6 K=2, N=3, duplicates
    cases.append((3, 3, [r[0], (r[0] * np.float32(1.0 + 1e-6)).astype(np.float32), r[2]],
                  [noisy(r[0], 0.25), noisy(r[0], 0.25), noisy(r[2], 0.2)]))                       # t7 near-duplicate refs
    out = []
    for K, N, refs, preds in cases:
        out.append((K, N, np.sum(refs, axis=0).astype(np.float32), refs, preds))
    return out


def _top2_gap(pair, K, N):
    """best - second-best mean SI-SDR over the DISTINCT assignments the reference enumerates (python floats)."""
    import itertools
    vals = {}
    for cols in itertools.combinations(range(N), K):
        for perm in itertools.permutations(range(K)):
            s = 0.0
            for r_ in range(K):
                s += float(pair[perm[r_]][cols[r_]])
            vals[tuple(cols[perm.index(i)] for i in range(K))] = s / float(K)
    v = sorted(vals.values(), reverse=True)
    return v[0] - v[1] if len(v) > 1 else float("inf")


def pit_many_cases():
    """More estimates than references (N = 4 .. 8; the reference enumerates combinations(range(N), K) for any N >= K,
    overlap3_core.py:81, evaluate_with_sources.py:352-355 / :421): clean and noisy permutations, duplicated
    estimates (ties between combinations), a silent estimate."""
    rng = np.random.default_rng(20261019)
    cases = []
    for ci, (K, N) in enumerate(((3, 4), (2, 4), (3, 5), (2, 5), (3, 6), (2, 7), (3, 8), (2, 8), (3, 4), (3, 7))):
        n = int(rng.integers(300, 4000))
        refs = [(rng.standard_normal(n) * 10 ** rng.uniform(-2.5, -1)).astype(np.float32) for _ in range(K)]
        extra = [(rng.standard_normal(n) * 0.03).astype(np.float32) for _ in range(N - K)]
        order = rng.permutation(N)
        noise = 10 ** rng.uniform(-2.0, -0.3)
        base = refs + extra
        preds = [(base[j] + noise * rng.standard_normal(n).astype(np.float32) * float(np.std(base[j]) + 1e-3)).astype(np.float32)
                 for j in order]
        if ci == 8:
            preds[3] = preds[0].copy()                   # duplicated estimate: two combinations tie exactly
        if ci == 9:
            preds[5] = np.zeros(n, np.float32)           # a silent estimate
        cases.append((K, N, np.sum(refs, axis=0).astype(np.float32), refs, preds))
    return cases


def make_pit_tie_golden(cases=None, fname="pit_golden_ties.npz"):
    o3, ews = shim.import_reference_pit()
    out = {}
    cases = pit_tie_cases() if cases is None else cases
    out["n_cases"] = np.int64(len(cases))
    for ci, (K, N, mix, refs, preds) in enumerate(cases):
        out[f"c{ci}_KN"] = np.array([K, N], np.int64)
        out[f"c{ci}_mix"] = mix
        for k, r in enumerate(refs):
            out[f"c{ci}_ref{k}"] = r
        for j, p in enumerate(preds):
            out[f"c{ci}_pred{j}"] = p
        pair = np.array([[o3._si_sdr(r, p) for p in preds] for r in refs], np.float64)
        best, sdri, idx = o3._compute_sdr_improvement_pit_k(mix, refs, preds)
        out[f"c{ci}_o3_pair"] = pair
        out[f"c{ci}_o3_res"] = np.array([best, sdri], np.float64)
        out[f"c{ci}_o3_idx"] = np.array(idx, np.int64)
        out[f"c{ci}_o3_gap"] = np.float64(_top2_gap(pair, K, N))
        if K == 2:
            best, sdri, idx, sw = ews.compute_sdr_improvement_pit(mix, refs[0], refs[1], preds)
            out[f"c{ci}_ews_swapped"] = np.int64(bool(sw))
        else:
            best, sdri, idx = ews.compute_sdr_improvement_pit_k(mix, refs, preds)
        out[f"c{ci}_ews_res"] = np.array([best, sdri], np.float64)
        out[f"c{ci}_ews_idx"] = np.array(idx, np.int64)
    np.savez_compressed(os.path.join(OUT, fname), **out)
    print("wrote", fname, "with", len(cases), "cases")


def make_separator_golden():
    from b200sep import synth, weights
    RefSeparator = shim.import_reference_separator()
    out = {}
    for n_src, lens in ((3, (6000, 1900)), (2, (4000,))):
        sd, args = weights.random_state_dict(n_src, seed=0)
        out[f"s{n_src}_digest"] = np.array(state_dict_digest(sd))

        def factory(sd=sd, args=args):
            m = ar.ConvTasNet(**args)
            m.load_state_dict(sd)
            return m

        if n_src == 3:
            # the reference's n_src == 3 branch: BaseModel.from_pretrained(<hub id>) (separation.py:43-47)
            ar.BaseModel.register_pretrained("JorisCos/ConvTasNet_Libri3Mix_sepclean_16k", factory)
            sep = RefSeparator(device="cpu", n_src=3, sample_rate=8000)
        else:
            # the n_src == 2 branch: ConvTasNet(n_src=2) + torch.load + load_state_dict (separation.py:50-56)
            path = os.path.join("/tmp", "b200sep_golden_ckpt2.pt")
            torch.save({"state_dict": sd}, path)
            sep = RefSeparator(device="cpu", n_src=2, sample_rate=8000, checkpoint=path)
        for i, n in enumerate(lens):
            mix, _ = synth.make_mixture(900 + 10 * n_src + i, n, n_src)
            outs = sep.separate(mix, 8000)
            out[f"s{n_src}_mix{i}"] = mix
            out[f"s{n_src}_out{i}"] = np.stack(outs).astype(np.float32)
        out[f"s{n_src}_n"] = np.int64(len(lens))
    np.savez_compressed(os.path.join(OUT, "separator_golden.npz"), **out)
    print("wrote separator_golden.npz")


if __name__ == "__main__":
    if not shim.reference_available():
        raise SystemExit("/root/reference not present: golden vectors can only be generated in the build container")
    os.makedirs(OUT, exist_ok=True)
    if "--many-only" in sys.argv:
        make_pit_tie_golden(pit_many_cases(), "pit_golden_many.npz")
        raise SystemExit(0)
    if "--ties-only" not in sys.argv:   # (the round-1 fixtures stay byte-identical unless regenerated on purpose)
        make_pit_golden()
        make_separator_golden()
    make_pit_tie_golden()
    make_pit_tie_golden(pit_many_cases(), "pit_golden_many.npz")
