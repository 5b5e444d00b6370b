"""Finer host-side breakdown of SeparationEngine.separate_batch (host entry point) on configs[1] (run on a B200)."""
import os, sys, time, ctypes as C
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import b200sep
from b200sep import synth, weights, _native as nat
sd, args = weights.random_state_dict(3, 0)
eng = b200sep.SeparationEngine(sd, args, 0, "fp16")
mixes = synth.make_batch(64, seed=2, n_src=3, sr=8000, lo_s=0.2, hi_s=4.0)
for _ in range(3):
    eng.separate_batch(mixes)
S = 3
dev = torch.device("cuda", 0)
for rep in range(4):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    segs = [np.ascontiguousarray(np.asarray(c).reshape(-1), dtype=np.float32) for c in mixes]
    lens = [int(a.shape[0]) for a in segs]
    total = int(sum(lens))
    arr = nat.i32_array(lens)
    ptrs = (C.c_void_p * len(segs))(*[a.ctypes.data for a in segs])
    t1 = time.perf_counter()
    need = int(nat.lib().b200sep_host_workspace_bytes(eng._handle, arr, len(lens)))
    t2 = time.perf_counter()
    pinned_out = torch.empty(total * S, dtype=torch.float32, pin_memory=True)
    t3 = time.perf_counter()
    ws_ptr = (eng._ws_host.data_ptr() + 255) // 256 * 256
    ws_bytes = eng._ws_host.numel() - (ws_ptr - eng._ws_host.data_ptr())
    stream = torch.cuda.current_stream(dev)
    nat.check(nat.lib().b200sep_separate_batch_host(eng._handle, ptrs, arr, len(lens), C.c_void_p(pinned_out.data_ptr()),
                                                    C.c_void_p(ws_ptr), C.c_size_t(ws_bytes), C.c_void_p(stream.cuda_stream), None), "x")
    t4 = time.perf_counter()
    stream.synchronize()
    t5 = time.perf_counter()
    t6 = time.perf_counter()
    hout = pinned_out.numpy(); res = []; o = 0
    for n in lens:
        seg = hout[o * S:(o + n) * S].reshape(S, n); res.append([seg[s] for s in range(S)]); o += n
    t7 = time.perf_counter()
    print(f"prep {1e3*(t1-t0):.2f} wsbytes {1e3*(t2-t1):.2f} pinned_alloc {1e3*(t3-t2):.2f} c_call {1e3*(t4-t3):.2f} "
          f"sync {1e3*(t5-t4):.2f} views {1e3*(t7-t6):.2f} total {1e3*(t7-t0):.2f} ms", flush=True)
    if rep % 2 == 0:
        del res, hout, pinned_out
t0 = time.perf_counter()
for _ in range(10): outs = eng.separate_batch(mixes)
print("separate_batch (keeping outs) avg ms", 1e3 * (time.perf_counter() - t0) / 10)
t0 = time.perf_counter()
for _ in range(10): eng.separate_batch(mixes)
print("separate_batch (dropping outs) avg ms", 1e3 * (time.perf_counter() - t0) / 10)
