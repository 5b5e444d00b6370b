"""Small end-to-end run for compute-sanitizer (memcheck): a ragged 5-segment batch, single calls, PIT, aux kernels."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import b200sep
from b200sep import aux, synth, weights
sd, args = weights.random_state_dict(3, 0)
sep = b200sep.Separator(backend="b200", device="cuda:0", sample_rate=8000, n_src=3, state_dict=sd, model_args=args)
mixes, srcs = synth.make_batch(5, seed=31, n_src=3, lo_s=0.2, hi_s=1.2, with_sources=True)
outs = sep.separate_batch(mixes, 8000)
one = sep.separate(mixes[0], 8000)
res = b200sep.pit_sisdr_batch(mixes, srcs, outs, "overlap3")
r2, _ = sep.engine.separate_and_score(mixes, srcs, "overlap3")
y = aux.resample_batch([mixes[0]], 16000, 8000, "linear")[0]
z = aux.resample_batch([mixes[0]], 16000, 8000, "sinc")[0]
m = aux.build_gt_overlap_mask(srcs[0][0], srcs[0][1], 8000, 0.025, 0.010, 0.05)
torch.cuda.synchronize()
print("ok", float(np.abs(outs[0][0]).mean()), res.perm.tolist(), y.shape, z.shape, int(m.sum()))
