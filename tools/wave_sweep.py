"""Timing sweep over the L2-wave size (device resident, configs[1] batch)."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import b200sep
from b200sep import synth, weights
sd, args = weights.random_state_dict(3, 0)
eng = b200sep.SeparationEngine(sd, args, 0, "fp16")
lens = sorted(synth.segment_lengths(64, 2)) if "sorted" in sys.argv else synth.segment_lengths(64, 2)
total = sum(lens)
d_mix = torch.randn(total, device="cuda") * 0.05
ws = torch.empty(eng.workspace_bytes(lens) + 256, dtype=torch.uint8, device="cuda")
out = torch.empty(total * 3, device="cuda")
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
ref = None
for wf in (0, 80000, 60000, 40000, 30000, 24000, 16000, 12000, 8000):
    for _ in range(3):
        eng.separate_device(d_mix, lens, out=out, workspace=ws, wave_frames=wf)
    torch.cuda.synchronize()
    ts = []
    for _ in range(5):
        flush.zero_()
        e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
        e0.record(); eng.separate_device(d_mix, lens, out=out, workspace=ws, wave_frames=wf); e1.record()
        torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    o = out.clone()
    if ref is None: ref = o
    ms = sum(ts) / len(ts)
    print(f"wave_frames={wf:6d} waves={len(eng.plan_waves(lens, wf)):2d}  {ms:7.3f} ms  {total/8000/ms*1e3:8.0f} audio-s/s  identical={bool(torch.equal(o, ref))}", flush=True)
