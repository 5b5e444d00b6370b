"""Host-side breakdown of Separator.separate_batch on configs[1] (run on a B200)."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import b200sep
from b200sep import synth, weights
sd, args = weights.random_state_dict(3, 0)
eng = b200sep.SeparationEngine(sd, args, 0, "fp16")
lens = synth.segment_lengths(64, 2)
mixes = [synth.make_mixture(i, n, 3)[0] for i, n in enumerate(lens)]
total = sum(lens)
for _ in range(3):
    eng.separate_batch(mixes)
def t(): torch.cuda.synchronize(); return time.perf_counter()
for rep in range(3):
    t0 = t()
    pin = torch.empty(total, dtype=torch.float32, pin_memory=True); h = pin.numpy(); o = 0
    for c, n in zip(mixes, lens):
        h[o:o + n] = c; o += n
    t1 = t()
    d_mix = torch.empty(total, dtype=torch.float32, device="cuda"); d_mix.copy_(pin, non_blocking=True)
    t2 = t()
    d_out = eng.separate_device(d_mix, lens)
    t2b = time.perf_counter()
    t3 = t()
    po = torch.empty(total * 3, dtype=torch.float32, pin_memory=True); po.copy_(d_out, non_blocking=True)
    t4 = t()
    hout = po.numpy(); res = []; o = 0
    for n in lens:
        seg = hout[o * 3:(o + n) * 3].reshape(3, n); res.append([seg[s] for s in range(3)]); o += n
    t5 = t()
    print(f"gather {1e3*(t1-t0):.2f}  h2d {1e3*(t2-t1):.2f}  forward {1e3*(t3-t2):.2f} (host issue {1e3*(t2b-t2):.2f})  d2h {1e3*(t4-t3):.2f}  views {1e3*(t5-t4):.2f}  total {1e3*(t5-t0):.2f} ms")
t0 = t()
for _ in range(5): eng.separate_batch(mixes)
print("separate_batch avg ms", 1e3 * (t() - t0) / 5)
