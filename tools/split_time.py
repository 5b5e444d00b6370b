"""Device-resident step time of configs[1] for the current B200SEP_SPLIT (+ output checksum): python tests/split_time.py"""
import os, sys, hashlib
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import b200sep
from b200sep import synth, weights
sd, args = weights.random_state_dict(3, 0)
eng = b200sep.SeparationEngine(sd, args, 0, "fp16")
nseg = int(sys.argv[1]) if len(sys.argv) > 1 else 64
lens = synth.segment_lengths(nseg, 2)
total = sum(lens)
g = torch.Generator(device="cuda").manual_seed(0)
d_mix = torch.randn(total, device="cuda", generator=g) * 0.05
ws = torch.empty(eng.workspace_bytes(lens) + 256, dtype=torch.uint8, device="cuda")
out = torch.empty(total * 3, device="cuda")
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for _ in range(3):
    eng.separate_device(d_mix, lens, out=out, workspace=ws)
torch.cuda.synchronize()
ts = []
for _ in range(10):
    flush.zero_()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record(); eng.separate_device(d_mix, lens, out=out, workspace=ws); e1.record()
    torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
h = hashlib.sha1(out.cpu().numpy().tobytes()).hexdigest()[:12]
print(f"split={os.environ.get('B200SEP_SPLIT', 'default')} nseg={nseg} min {min(ts):.3f} ms  mean {sum(ts)/len(ts):.3f} ms  "
      f"{total/8000/(sum(ts)/len(ts))*1e3:.0f} audio-s/s  sha={h}", flush=True)
