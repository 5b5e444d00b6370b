"""clock64 stamps of the mask+decoder kernel's roles (run on a B200): python tools/timeline_tail.py
slots: 0+c MMA1(c) issue, 16+c MMA2(c) issue, 32+c epilogue reaches chunk c, 48+c accumulator ready, 64+c staged, 80+c weight TMA issue."""
import os, sys, ctypes as C
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import b200sep
from b200sep import _native as nat, synth, weights
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
eng.separate_device(d_mix, lens, out=out, workspace=ws)
torch.cuda.synchronize()
nat.check(nat.lib().b200sep_debug_timeline(None), "timeline")
t = tl.cpu().numpy().reshape(148, 8, 128)
for cta in (0, 77):
    nz = t[cta][t[cta] > 0]
    t0 = nz.min()
    print("== CTA", cta, "span", nz.max() - t0)
    for it in (2, 3):
        row = t[cta, it]
        for name, lo in (("tma", 80), ("mma1", 0), ("ep_in", 32), ("acc_ok", 48), ("staged", 64), ("mma2", 16)):
            print(f"  tile {it} {name:7s}", " ".join(str(int(row[lo + c] - t0)) for c in range(12)))
