"""Checks the tensor-memory-A MMA (single CTA and CTA pair) against a plain GEMM of the fp16-rounded operands."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from b200sep import _native as nat
g = torch.Generator().manual_seed(0)
w = torch.randn(256, 64, generator=g).cuda()
h = torch.randn(256, 64, generator=g).cuda()
ref = w.half().float() @ h.half().float().t()
for pair in (0, 1):
    out = torch.full((256, 256), float("nan"), device="cuda")
    rc = nat.lib().b200sep_debug_ts_probe(C.c_void_p(w.data_ptr()), C.c_void_p(h.data_ptr()), C.c_void_p(out.data_ptr()), pair, None)
    if rc:
        print("pair", pair, "FAILED:", nat.last_error()); continue
    n = 256 if pair else 128
    err = (out[:n, :n] - ref[:n, :n]).abs().max().item()
    print(f"pair={pair}: max abs err {err:.3e} (ref absmax {ref.abs().max().item():.2f})", flush=True)
    if err > 1e-2:
        d = (out[:n, :n] - ref[:n, :n]).abs()
        bad = (d > 1e-2)
        print("  bad rows", bad.any(1).nonzero().flatten()[:16].tolist(), "bad cols", bad.any(0).nonzero().flatten()[:16].tolist())
        print("  out[0,:8]", out[0, :8].tolist(), "\n  ref[0,:8]", ref[0, :8].tolist())
