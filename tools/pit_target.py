"""Profiling / timing target of the PIT kernel at BASELINE.json configs[3] size: 3000 segments U[0.4, 4] s, K = N = 3."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import b200sep
from b200sep import synth
from b200sep.pit import pit_sisdr_device
lens = synth.segment_lengths(3000, 4, 0.4, 4.0)
total = sum(lens)
g = torch.Generator(device="cuda").manual_seed(0)
d_mix = torch.randn(total, device="cuda", generator=g) * 0.05
d_ref = torch.randn(3 * total, device="cuda", generator=g) * 0.05
d_est = d_ref + 0.01 * torch.randn(3 * total, device="cuda", generator=g)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for _ in range(3):
    pit_sisdr_device(d_ref, d_est, d_mix, lens, 3, 3, "overlap3")
torch.cuda.synchronize()
ts = []
for _ in range(int(sys.argv[1]) if len(sys.argv) > 1 else 5):
    flush.zero_()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record(); pit_sisdr_device(d_ref, d_est, d_mix, lens, 3, 3, "overlap3"); e1.record()
    torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
ms = min(ts)
print(f"pit 3000 segments, {total} samples: {ms:.3f} ms per call (host marshalling + two launches, GPU idle in between) = "
      f"{28 * total / ms / 1e6:.1f} GB/s (28 B/sample algorithmic), {total / 8000 / ms * 1e3:.0f} audio-s/s", flush=True)
# kernel-level: CUDA events around the two launches inside the C entry point
import ctypes as C
from b200sep import _native as nat
nat.lib().b200sep_pit_profile_enable(1)
km = []
for _ in range(5):
    flush.zero_()
    pit_sisdr_device(d_ref, d_est, d_mix, lens, 3, 3, "overlap3")
    out = (C.c_float * 2)()
    nat.check(nat.lib().b200sep_pit_profile_read(out), "pit_profile_read")
    km.append((out[0], out[1]))
nat.lib().b200sep_pit_profile_enable(0)
m0 = min(k[0] for k in km); f0 = min(k[1] for k in km)
print(f"   kernels: moments {m0 * 1e3:.1f} us = {28 * total / m0 / 1e6:.1f} GB/s, finalize {f0 * 1e3:.1f} us", flush=True)
