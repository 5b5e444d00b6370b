"""Energy per launch of the forward's kernels: run the forward truncated after k launches (b200sep_debug_set_stop) in a
loop for ~1.5 s, sample NVML power / SM clock, report time and energy per iteration; differences between consecutive k
give one kernel's energy.  python tools/power_probe.py [pair|single]"""
import os, sys, time, threading
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pynvml
import b200sep
from b200sep import _native as nat, synth, weights

pynvml.nvmlInit()
h = pynvml.nvmlDeviceGetHandleByIndex(0)
sd, args = weights.random_state_dict(3, 0)
eng = b200sep.SeparationEngine(sd, args, 0, "fp16")
lens = synth.segment_lengths(64, 2)
d_mix = torch.randn(sum(lens), device="cuda") * 0.05
ws = torch.empty(eng.workspace_bytes(lens) + 256, dtype=torch.uint8, device="cuda")
out = torch.empty(sum(lens) * 3, device="cuda")
os.environ.setdefault("B200SEP_SPLIT", "1")

def sample(stop_evt, acc):
    while not stop_evt.is_set():
        acc.append((pynvml.nvmlDeviceGetPowerUsage(h) / 1e3, pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)))
        time.sleep(0.02)

prev = None
for stop in [int(v) for v in (sys.argv[1:] or ["2", "3", "4", "17", "18", "52"])]:
    nat.lib().b200sep_debug_set_stop(eng._handle, stop % 52)
    for _ in range(5):
        eng.separate_device(d_mix, lens, out=out, workspace=ws)
    torch.cuda.synchronize()
    acc, ev = [], threading.Event()
    th = threading.Thread(target=sample, args=(ev, acc)); th.start()
    t0 = time.perf_counter(); n = 0
    while time.perf_counter() - t0 < 1.5:
        for _ in range(20):
            eng.separate_device(d_mix, lens, out=out, workspace=ws)
        torch.cuda.synchronize(); n += 20
    dt = (time.perf_counter() - t0) / n
    ev.set(); th.join()
    late = acc[len(acc) // 2:]
    pw = float(np.mean([a[0] for a in late])); ck = float(np.median([a[1] for a in late]))
    e = pw * dt * 1e3
    msg = f"stop={stop:3d}: {dt * 1e3:7.3f} ms/iter  {pw:6.0f} W  {ck:5.0f} MHz  {e:7.2f} mJ/iter"
    if prev is not None:
        msg += f"   delta: {(dt - prev[0]) * 1e6:7.1f} us  {e - prev[1]:7.2f} mJ  -> {(e - prev[1]) / max(dt - prev[0], 1e-9) / 1e3:6.0f} W marginal"
    print(msg, flush=True)
    prev = (dt, e)
