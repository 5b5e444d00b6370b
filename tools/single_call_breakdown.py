"""Where one separate() call of a single segment spends its time (run on a B200):
per-kernel-class CUDA-event times of the 52 launches run one after the other (profiling serialises them), the
device time of the graph replay (PDL edges, launches overlapped), and the host call around it."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import b200sep
from b200sep import synth, weights
from b200sep import _native as nat
from bench import classify_launches, KERNEL_CLASSES

sd, args = weights.random_state_dict(3, 0)
eng = b200sep.SeparationEngine(sd, args, 0, "fp16")
pipeline = int(nat.lib().b200sep_pipeline(eng._handle))
for sec in (0.4, 1.0, 4.0):
    n = int(sec * 8000)
    lens = [n]
    mix, _ = synth.make_mixture(7, n, 3)
    d_mix = torch.from_numpy(mix).cuda()
    ws = torch.empty(eng.workspace_bytes(lens) + 256, dtype=torch.uint8, device="cuda")
    out = torch.empty(n * 3, device="cuda")
    for _ in range(5):
        eng.separate_device(d_mix, lens, out=out, workspace=ws)
    torch.cuda.synchronize()
    ts = []
    for _ in range(20):
        e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
        e0.record(); eng.separate_device(d_mix, lens, out=out, workspace=ws); e1.record()
        torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    t0 = time.perf_counter()
    for _ in range(50):
        eng.separate_batch([mix])
    host = (time.perf_counter() - t0) / 50
    eng.profile_last_call(True)
    acc = {k: [] for k in KERNEL_CLASSES}
    for _ in range(3):
        eng.separate_device(d_mix, lens, out=out, workspace=ws)
        for k, v in classify_launches(eng.profile_read(), 24, pipeline).items():
            acc[k].extend(v)
    eng.profile_last_call(False)
    serial = sum(sum(v) / 3 for v in acc.values())
    print(f"{sec} s segment: device (graph replay) min {1e3*min(ts):.0f} us median {1e3*float(np.median(ts)):.0f} us | "
          f"separate_batch([mix]) host call {1e3*host:.3f} ms | serial sum of launches {1e3*serial:.0f} us", flush=True)
    print("    per launch us:", " ".join(f"{k}={1e3*sum(v)/len(v):.1f}" for k, v in acc.items() if v), flush=True)
