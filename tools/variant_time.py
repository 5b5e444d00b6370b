"""Per-kernel-class timing of one library variant (B200SEP_LIB=...): python tests/variant_time.py"""
import os, sys, json
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import b200sep
from b200sep import synth, weights
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import classify_launches, KERNEL_CLASSES
from b200sep import _native as nat
sd, args = weights.random_state_dict(3, 0)
eng = b200sep.SeparationEngine(sd, args, 0, "fp16")
pipeline = int(nat.lib().b200sep_pipeline(eng._handle))
lens = synth.segment_lengths(64, 2)
total = sum(lens)
d_mix = torch.randn(total, device="cuda") * 0.05
ws = torch.empty(eng.workspace_bytes(lens) + 256, dtype=torch.uint8, device="cuda")
out = torch.empty(total * 3, device="cuda")
for _ in range(3):
    eng.separate_device(d_mix, lens, out=out, workspace=ws)
torch.cuda.synchronize()
ts = []
for _ in range(5):
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record(); eng.separate_device(d_mix, lens, out=out, workspace=ws); e1.record()
    torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
eng.profile_last_call(True)
acc = {k: [] for k in KERNEL_CLASSES}
for _ in range(3):
    eng.separate_device(d_mix, lens, out=out, workspace=ws)
    for k, v in classify_launches(eng.profile_read(), 24, pipeline).items():
        acc[k].extend(v)
print(os.path.basename(os.environ.get("B200SEP_LIB", "default")), os.environ.get("B200SEP_KA", "pair"), f"{min(ts):.3f} ms",
      " ".join(f"{k}={1e3*sum(v)/len(v):.1f}" for k, v in acc.items() if v), flush=True)
ka = acc["dw_rs_gemm"]
if ka:  # per dilation (block i has d = 2^(i % 8)): mean over the 3 repeats x 3 stack repeats
    n = len(ka) // 24
    print("   K_A us by dilation:", " ".join(f"d{1 << k}={1e3 * np.mean([ka[r * 24 + b] for r in range(n) for b in range(24) if b % 8 == k]):.1f}" for k in range(8)), flush=True)
kb = acc["apply_in_gemm"]
if kb:
    print(f"   K_B us: block0={1e3 * np.mean(kb[0::24]):.1f} others={1e3 * np.mean([v for i, v in enumerate(kb) if i % 24]):.1f}", flush=True)
