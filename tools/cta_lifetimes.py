"""Wall-clock (globaltimer) lifetimes of the CTAs of one K_B and one K_A launch (run on a B200):
python tests/cta_lifetimes.py [stop_after]  -- stop_after = launches to run (default 10: the K_A of block 3).
Needs a library built with -DB200SEP_TL_NS=1 (scripts_gpu/build_variant.sh TLNS -DB200SEP_TL_NS=1; B200SEP_LIB=...)."""
import os, sys, ctypes as C
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import b200sep
from b200sep import _native as nat, synth, weights
stop = int(sys.argv[1]) if len(sys.argv) > 1 else 10
sd, args = weights.random_state_dict(3, 0)
eng = b200sep.SeparationEngine(sd, args, 0, "fp16")
lens = synth.segment_lengths(64, 2)
d_mix = torch.randn(sum(lens), device="cuda") * 0.05
ws = torch.empty(eng.workspace_bytes(lens) + 256, dtype=torch.uint8, device="cuda")
out = torch.empty(sum(lens) * 3, device="cuda")
for _ in range(2):
    eng.separate_device(d_mix, lens, out=out, workspace=ws)
torch.cuda.synchronize()
tl = torch.zeros(148 * 8 * 128, dtype=torch.int64, device="cuda")
nat.check(nat.lib().b200sep_debug_timeline(C.c_void_p(tl.data_ptr())), "timeline")
nat.lib().b200sep_debug_set_stop(eng._handle, stop)
e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
e0.record(); eng.separate_device(d_mix, lens, out=out, workspace=ws); e1.record()
torch.cuda.synchronize()
nat.check(nat.lib().b200sep_debug_timeline(None), "timeline")
t = tl.cpu().numpy().reshape(148, 8, 128)
for name, a, b in (("K_B", 120, 121), ("K_A", 124, 125)):
    st, en = t[:, 7, a].astype(np.float64), t[:, 7, b].astype(np.float64)
    ok = (st > 0) & (en > 0)
    if not ok.any():
        continue
    t0 = st[ok].min()
    life = (en[ok] - st[ok]) / 1e3
    print(f"{name}: {ok.sum()} CTAs; first start -> last end {(en[ok].max() - t0) / 1e3:.1f} us; start skew {(st[ok].max() - t0) / 1e3:.1f} us; "
          f"CTA lifetime min/median/max {life.min():.1f}/{np.median(life):.1f}/{life.max():.1f} us; end spread {(en[ok].max() - en[ok].min()) / 1e3:.1f} us")
    # clock64 span inside the same CTAs (first tile stamp -> last) for comparison
    lo, hi = (0, 32) if name == "K_B" else (32, 120)
    sub = t[:, :7, lo:hi].reshape(148, -1).astype(np.float64)
    sub[sub <= 0] = np.nan
    span = np.nanmax(sub, axis=1) - np.nanmin(sub, axis=1)
    print(f"      clock64 span of the stamped pipeline min/median/max {np.nanmin(span) / 1e3:.1f}/{np.nanmedian(span) / 1e3:.1f}/{np.nanmax(span) / 1e3:.1f} k cycles")
