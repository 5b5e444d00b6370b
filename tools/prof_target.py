"""Profiling target: N forwards of BASELINE.json configs[1] (64 segments) on cuda:0, device resident."""
import os, sys
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import b200sep
from b200sep import synth, weights

n_iter = int(sys.argv[1]) if len(sys.argv) > 1 else 2
prec = sys.argv[2] if len(sys.argv) > 2 else "fp16"
sd, args = weights.random_state_dict(3, 0)
eng = b200sep.SeparationEngine(sd, args, 0, prec)
lens = synth.segment_lengths(64, 2)
d_mix = torch.randn(sum(lens), device="cuda") * 0.05
ws = torch.empty(eng.workspace_bytes(lens) + 256, dtype=torch.uint8, device="cuda")
out = torch.empty(sum(lens) * 3, device="cuda")
for _ in range(n_iter):
    eng.separate_device(d_mix, lens, out=out, workspace=ws)
torch.cuda.synchronize()
print("ok", float(out.abs().mean()))
