"""Latency of the reference-shaped call: one Separator.separate(chunk, sr) per overlap segment (run on a B200)."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import b200sep
from b200sep import synth, weights
sd, args = weights.random_state_dict(3, 0)
sep = b200sep.Separator(backend="b200", device="cuda:0", sample_rate=8000, n_src=3, state_dict=sd, model_args=args)
for dur in (0.4, 1.0, 2.0, 4.0):
    mix, _ = synth.make_mixture(7, int(dur * 8000), 3)
    for _ in range(5):
        sep.separate(mix, 8000)
    t0 = time.perf_counter()
    n = 50
    for _ in range(n):
        out = sep.separate(mix, 8000)
    dt = (time.perf_counter() - t0) / n
    print(f"{dur:.1f} s segment: {1e3 * dt:.3f} ms per separate() call = {dur / dt:.0f} audio-s/s", flush=True)
