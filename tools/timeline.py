"""In-kernel timeline of the fused block kernels (run on a B200): python tests/timeline.py [stop_after]
Prints, for a few CTAs, clock64 stamps (in cycles relative to the CTA's first stamp) of every pipeline role."""
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
eng.separate_device(d_mix, lens, out=out, workspace=ws)
torch.cuda.synchronize()
nat.check(nat.lib().b200sep_debug_timeline(None), "timeline")
t = tl.cpu().numpy().reshape(148, 8, 128)
np.save("gpurun_out/timeline.npy", t)
for cta in (0, 77, 147):
    print("== CTA", cta)
    for name, lo, hi in (("K_B", 0, 32), ("K_A", 32, 128)):
        sub = t[cta, :, lo:hi]
        nz = sub[sub > 0]
        if nz.size == 0:
            continue
        t0 = nz.min()
        print(f" {name}: span {nz.max() - t0} cycles")
        for it in range(8):
            row = sub[it]
            if not (row > 0).any():
                continue
            print(f"  tile {it}: " + " ".join(f"{lo + k}:{int(v - t0)}" for k, v in enumerate(row) if v > 0))
