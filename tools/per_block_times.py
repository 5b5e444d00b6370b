"""Per-TCN-block kernel times of the fused pipeline (K_B, K_A by dilation): python tests/per_block_times.py"""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import b200sep
from b200sep import synth, weights
sd, args = weights.random_state_dict(3, 0)
eng = b200sep.SeparationEngine(sd, args, 0, "fp16")
lens = synth.segment_lengths(64, 2)
total = sum(lens)
d_mix = torch.randn(total, device="cuda") * 0.05
ws = torch.empty(eng.workspace_bytes(lens) + 256, dtype=torch.uint8, device="cuda")
out = torch.empty(total * 3, device="cuda")
for _ in range(3):
    eng.separate_device(d_mix, lens, out=out, workspace=ws)
eng.profile_last_call(True)
acc = []
for _ in range(4):
    eng.separate_device(d_mix, lens, out=out, workspace=ws)
    acc.append(eng.profile_read())
ms = np.mean(np.array(acc), axis=0) * 1e3
print("enc %.1f bneck %.1f" % (ms[0], ms[1]))
for i in range(24):
    print("block %2d d=%3d  K_B %.1f  K_A %.1f" % (i, 1 << (i % 8), ms[2 + 2 * i], ms[3 + 2 * i]))
print("tail", ms[50:])
