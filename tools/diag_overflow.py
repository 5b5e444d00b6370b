import os, sys, warnings
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["B200SEP_CONDITION"] = "0"
os.environ["B200SEP_FP32_FALLBACK"] = "0"
import numpy as np, torch
import b200sep
from b200sep import weights, synth, _native as nat
sd, args = weights.random_state_dict(3, 0)
for s in (1e4, 1e6):
    sd2 = {k: v.clone() for k, v in sd.items()}
    for k in sd2:
        if k.endswith(("shared_block.0.weight", "shared_block.0.bias", "shared_block.3.weight", "shared_block.3.bias")):
            sd2[k] = sd2[k] * s
    sd2["encoder.filterbank._filters"] = sd2["encoder.filterbank._filters"] * s
    eng = b200sep.SeparationEngine(sd2, args, 0, "fp16")
    m = synth.make_mixture(1, 16000, 3)[0]
    with warnings.catch_warnings(record=True) as w:
        warnings.simplefilter("always")
        out = eng.separate_batch([m])
    o = np.stack(out[0])
    print("s", s, "finite frac", np.isfinite(o).mean(), "absmax", np.nanmax(np.abs(o)), "warnings", [str(x.message)[:60] for x in w], flush=True)
    eng.close()
