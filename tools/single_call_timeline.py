"""Kernel-level timeline of the LAST TCN block of a single-segment forward (run on a B200): entry / dependency wait
returned / exit of K_B and K_A in globaltimer ns, plus the role stamps of the (single) visit in cycles.
    python tools/single_call_timeline.py [seconds]"""
import os, sys, ctypes as C
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import b200sep
from b200sep import _native as nat, synth, weights

sec = float(sys.argv[1]) if len(sys.argv) > 1 else 4.0
sd, args = weights.random_state_dict(3, 0)
eng = b200sep.SeparationEngine(sd, args, 0, "fp16")
n = int(sec * 8000)
lens = [n]
d_mix = torch.randn(n, device="cuda") * 0.05
ws = torch.empty(eng.workspace_bytes(lens) + 256, dtype=torch.uint8, device="cuda")
out = torch.empty(n * 3, device="cuda")
for _ in range(3):
    eng.separate_device(d_mix, lens, out=out, workspace=ws)
torch.cuda.synchronize()
tl = torch.zeros(148 * 8 * 128, dtype=torch.int64, device="cuda")
nat.check(nat.lib().b200sep_debug_timeline(C.c_void_p(tl.data_ptr())), "timeline")
eng.separate_device(d_mix, lens, out=out, workspace=ws)
torch.cuda.synchronize()
nat.check(nat.lib().b200sep_debug_timeline(None), "timeline")
t = tl.cpu().numpy().reshape(148, 8, 128)
n_cta = int((t[:, 7, 24] > 0).sum())
print(f"{sec} s segment, {n_cta} CTAs stamped")
kb0 = t[:n_cta, 7, 24].min()
for name, base in (("K_B", 24), ("K_A", 112)):
    e, w, x = t[:n_cta, 7, base], t[:n_cta, 7, base + 1], t[:n_cta, 7, base + 2]
    ce, cw, cx = t[:n_cta, 7, base + 4], t[:n_cta, 7, base + 5], t[:n_cta, 7, base + 6]
    print(f"{name}: entry {int(e.min() - kb0)}..{int(e.max() - kb0)} ns | wait returned {int(w.min() - kb0)}..{int(w.max() - kb0)} ns | "
          f"exit {int(x.min() - kb0)}..{int(x.max() - kb0)} ns   (relative to the first K_B entry)")
    print(f"     per CTA, cycles: entry->wait {int(np.median(cw - ce))}  wait->exit {int(np.median(cx - cw))} (median)")
for cta in (0, n_cta - 1):
    for name, lo, hi, base in (("K_B", 0, 24, 24), ("K_A", 32, 112, 112)):
        row = t[cta, 0, lo:hi]
        c0 = t[cta, 7, base + 5]   # clock64 when the dependency wait returned
        print(f" CTA {cta} {name} visit 0, cycles after the wait returned: " +
              " ".join(f"{lo + k}:{int(v - c0)}" for k, v in enumerate(row) if v > 0) + f" exit:{int(t[cta, 7, base + 6] - c0)}")
