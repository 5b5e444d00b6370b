"""GPU bring-up diagnostics (run with `python tests/gpu_diag.py` on a B200; not a pytest file).

Prints, stage by stage, how far the CUDA path is from the oracle so a single GPU call localises a
bug: raw GEMM engines, then every intermediate activation of the first TCN block, then end-to-end
parity in the three precision modes, PIT, and a quick timing.
"""
from __future__ import annotations

import ctypes as C
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import b200sep  # noqa: E402
from b200sep import _native as nat, synth, weights  # noqa: E402
from oracle import asteroid_restated as ar  # noqa: E402
from oracle import pit_restated as pr  # noqa: E402

dev = torch.device("cuda:0")


def db(ref: np.ndarray, est: np.ndarray) -> float:
    ref = np.asarray(ref, np.float64).ravel()
    est = np.asarray(est, np.float64).ravel()
    err = float(np.sum((ref - est) ** 2)) + 1e-300
    return 10 * np.log10((float(np.sum(ref ** 2)) + 1e-300) / err)


def gemm_check():
    print("== raw GEMM engines (fp16 operands): SNR of C vs torch fp32 matmul of the rounded operands")
    L = nat.lib()
    g = torch.Generator(device="cpu").manual_seed(0)
    for (bn, K, N, M) in [(256, 128, 512, 256), (256, 512, 256, 384), (128, 512, 128, 256), (128, 128, 256, 128),
                          (256, 128, 1536, 128 * 40)]:
        a = torch.randn(M, K, generator=g).to(dev)
        b = torch.randn(N, K, generator=g).to(dev)
        ref = (a.half().float() @ b.half().float().t()).cpu().numpy()
        for use_tc in (0, 1):
            c = torch.full((M, N), float("nan"), device=dev)
            rc = L.b200sep_debug_gemm(C.c_void_p(a.data_ptr()), C.c_void_p(b.data_ptr()), C.c_void_p(c.data_ptr()),
                                      M, N, K, bn, use_tc, None)
            if rc:
                print(f"  bn={bn} K={K} N={N} M={M} tc={use_tc}: ERROR {nat.last_error()}")
                continue
            torch.cuda.synchronize()
            cc = c.cpu().numpy()
            print(f"  bn={bn} K={K} N={N} M={M} tc={use_tc}: SNR {db(ref, cc):7.1f} dB  nan={int(np.isnan(cc).sum())}")


def tile_T_to_CF(buf: torch.Tensor, Cn: int, tile0: int, ntiles: int, F: int) -> np.ndarray:
    """[nt][C][128] -> [C, F] of one segment."""
    v = buf.view(-1, Cn, 128)[tile0:tile0 + ntiles]  # [ntiles, C, 128]
    return v.permute(1, 0, 2).reshape(Cn, ntiles * 128)[:, :F].float().cpu().numpy()


def oracle_intermediates(model, wav: np.ndarray):
    """tf_rep, x after bottleneck, then per block (h1, h2, x, skip_sum)."""
    out = {}
    with torch.no_grad():
        w = torch.from_numpy(wav)[None, None]
        tf = model.encoder(w)
        out["enc"] = tf[0].numpy()
        x = model.masker.bottleneck(tf)
        out["x_bn"] = x[0].numpy()
        skip = 0
        for i, blk in enumerate(model.masker.TCN):
            sb = blk.shared_block
            h1 = sb[1](sb[0](x))
            h2 = sb[4](sb[3](sb[2](h1)))
            y = sb[5](h2)
            res, sk = blk.res_conv(y), blk.skip_conv(y)
            skip = skip + sk
            x = x + res
            if i < 2:
                out[f"h1_{i}"] = h1[0].numpy()
                out[f"h2_{i}"] = h2[0].numpy()
                out[f"x_{i}"] = x[0].numpy()
                out[f"skip_{i}"] = skip[0].numpy()
    return out


def stage_check(precision: str):
    print(f"== stage-by-stage, precision={precision}")
    sd, args = weights.random_state_dict(3, 0)
    model = ar.ConvTasNet(**args).eval()
    model.load_state_dict(sd)
    lens = [3000, 1700]
    mixes = [synth.make_mixture(7 + i, n, 3)[0] for i, n in enumerate(lens)]
    eng = b200sep.SeparationEngine(sd, args, 0, precision)
    lay = (C.c_uint64 * 16)()
    nat.check(nat.lib().b200sep_debug_layout(eng._handle, nat.i32_array(lens), len(lens), lay), "layout")
    o_enc, o_x, o_xa, o_skip, o_skipa, o_h1, o_h2, o_masked, n_tiles, total = [int(v) for v in lay][:10]
    half = precision != "fp32"
    es = 2 if half else 4
    tdt = torch.float16 if half else torch.float32
    ws = torch.zeros(total + 256, dtype=torch.uint8, device=dev)
    base = (ws.data_ptr() + 255) // 256 * 256 - ws.data_ptr()
    d_mix = torch.from_numpy(np.concatenate(mixes)).to(dev)
    Fs = [(n - 16) // 8 + 1 for n in lens]
    nts = [(f + 127) // 128 for f in Fs]
    t0s = [0, nts[0]]
    refs = [oracle_intermediates(model, m) for m in mixes]

    def view(off, Cn, dtype, esz):
        n = n_tiles * Cn * 128
        return ws[base + off: base + off + n * esz].view(dtype)

    def run(stop):
        nat.lib().b200sep_debug_set_stop(eng._handle, stop)
        ws.zero_()
        eng.separate_device(d_mix, lens, workspace=ws)
        torch.cuda.synchronize()

    def report(name, off, Cn, dtype, esz, key):
        buf = view(off, Cn, dtype, esz)
        vals = [db(refs[s][key], tile_T_to_CF(buf, Cn, t0s[s], nts[s], Fs[s])) for s in range(2)]
        print(f"  {name:10s} vs oracle {key:8s}: " + "  ".join(f"{v:7.1f} dB" for v in vals))

    run(1); report("enc", o_enc, 512, tdt, es, "enc")
    run(2); report("x(bottlenk)", o_x, 128, torch.float32, 4, "x_bn")
    if nat.lib().b200sep_pipeline(eng._handle) == 1:
        # fused pipeline: launch 3+2i = K_B(i) (applies block i-1, writes h1[i]), 4+2i = K_A(i) (raw G)
        run(3); report("h1[0]", o_h1, 512, tdt, es, "h1_0")
        run(5); report("h1[1]", o_h1, 512, tdt, es, "h1_1"); report("x[0]", o_x, 128, torch.float32, 4, "x_0")
        run(7); report("x[1]", o_x, 128, torch.float32, 4, "x_1")
        nat.lib().b200sep_debug_set_stop(eng._handle, 0)
        eng.close()
        return
    for i in range(2):
        run(3 + 3 * i); report(f"h1[{i}]", o_h1, 512, tdt, es, f"h1_{i}")
        run(4 + 3 * i); report(f"h2[{i}]", o_h2, 512, tdt, es, f"h2_{i}")
        run(5 + 3 * i); report(f"x[{i}]", o_x, 128, torch.float32, 4, f"x_{i}"); report(f"skip[{i}]", o_skip, 128, torch.float32, 4, f"skip_{i}")
    nat.lib().b200sep_debug_set_stop(eng._handle, 0)
    eng.close()


def e2e_check():
    print("== end-to-end parity vs oracle fp32 CPU (min over sources/segments of SI-SDR, dB)")
    for n_src in (3, 2):
        sd, args = weights.random_state_dict(n_src, 1)
        model = ar.ConvTasNet(**args).eval()
        model.load_state_dict(sd)
        lens = [32000, 1600, 4321, 16, 23, 24, 1039, 1040, 1041, 9999]
        mixes = [synth.make_mixture(50 + i, n, n_src)[0] for i, n in enumerate(lens)]
        with torch.no_grad():
            refs = [model(torch.from_numpy(m)[None])[0].numpy() for m in mixes]
        for prec in ("fp32", "fp16_simt", "fp16"):
            try:
                sep = b200sep.Separator(device="cuda:0", sample_rate=8000, n_src=n_src, state_dict=sd,
                                        model_args=args, precision=prec)
                outs = sep.separate_batch(mixes, 8000)
                vals = [min(pr.si_sdr_f64(refs[i][s], outs[i][s]) for s in range(n_src)) for i in range(len(lens))]
                tails = [float(np.abs(outs[i][0][-(((lens[i] - 16) % 8)) or len(outs[i][0]):]).max()) if (lens[i] - 16) % 8 else 0.0
                         for i in range(len(lens))]
                print(f"  n_src={n_src} {prec:9s}: " + " ".join(f"{v:6.1f}" for v in vals) + f"   tail_max={max(tails):.1e}")
                sep.engine.close()
            except Exception as e:  # noqa: BLE001
                print(f"  n_src={n_src} {prec}: FAILED {type(e).__name__}: {e} / {e.__cause__}")


def pit_check():
    print("== PIT kernel vs oracle restatement")
    rng = np.random.default_rng(3)
    for (K, N) in [(3, 3), (2, 2), (2, 3)]:
        mixes, refs, preds = [], [], []
        for i in range(6):
            n = int(rng.integers(200, 40000))
            r = [rng.standard_normal(n).astype(np.float32) * 0.05 for _ in range(K)]
            order = rng.permutation(N)
            base = r + [rng.standard_normal(n).astype(np.float32) * 0.05 for _ in range(N - K)]
            p = [(base[j] + 0.3 * rng.standard_normal(n).astype(np.float32) * 0.05).astype(np.float32) for j in order]
            mixes.append(np.sum(r, axis=0).astype(np.float32)); refs.append(r); preds.append(p)
        for variant in ("overlap3", "evalsrc"):
            res = b200sep.pit_sisdr_batch(mixes, refs, preds, variant)
            ok, worst = True, 0.0
            for i in range(len(mixes)):
                if variant == "overlap3":
                    best, sdri, idx = pr.compute_sdr_improvement_pit_k_o3(mixes[i], refs[i], preds[i])
                elif K == 2:
                    best, sdri, idx, _sw = pr.compute_sdr_improvement_pit_ews(mixes[i], refs[i][0], refs[i][1], preds[i])
                else:
                    best, sdri, idx = pr.compute_sdr_improvement_pit_k_ews(mixes[i], refs[i], preds[i])
                ok &= list(res.perm[i]) == list(idx)
                worst = max(worst, abs(res.sisdr[i] - best), abs(res.sisdri[i] - sdri))
            print(f"  K={K} N={N} {variant:8s}: perms_equal={ok}  max|d dB|={worst:.2e}")


def quick_timing():
    print("== quick timing, 64 segments U[0.2,4] s, device resident")
    sd, args = weights.random_state_dict(3, 0)
    lens = synth.segment_lengths(64, 2)
    total = sum(lens)
    d_mix = (torch.randn(total, device=dev) * 0.05)
    for prec in ("fp16", "fp16_simt", "fp32"):
        try:
            eng = b200sep.SeparationEngine(sd, args, 0, prec)
            ws = torch.empty(eng.workspace_bytes(lens) + 256, dtype=torch.uint8, device=dev)
            out = torch.empty(total * 3, device=dev)
            for _ in range(2):
                eng.separate_device(d_mix, lens, out=out, workspace=ws)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
            e0.record()
            n_it = 5
            for _ in range(n_it):
                eng.separate_device(d_mix, lens, out=out, workspace=ws)
            e1.record(); torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / n_it
            print(f"  {prec:9s}: {ms:8.2f} ms / batch of {total / 8000:.1f} audio-s -> {total / 8000 / ms * 1e3:9.0f} audio-s/s"
                  f"  ({total / 8000 * 9.97e9 / ms / 1e9:.1f} TFLOP/s)")
            eng.close()
        except Exception as e:  # noqa: BLE001
            print(f"  {prec}: FAILED {e} / {e.__cause__}")


if __name__ == "__main__":
    which = sys.argv[1:] or ["gemm", "stage32", "stage16s", "stage16", "e2e", "pit", "time"]
    for w in which:
        t = time.time()
        try:
            {"gemm": gemm_check, "stage32": lambda: stage_check("fp32"), "stage16s": lambda: stage_check("fp16_simt"),
             "stage16": lambda: stage_check("fp16"), "e2e": e2e_check, "pit": pit_check, "time": quick_timing}[w]()
        except Exception as e:  # noqa: BLE001
            import traceback
            traceback.print_exc()
            print(f"!! {w} failed: {e}")
        print(f"   [{w}: {time.time() - t:.1f} s]", flush=True)
