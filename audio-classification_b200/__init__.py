"""b200sep -- B200-native Conv-TasNet separation + PIT SI-SDR backend (sm_100a CUDA behind a C ABI).

Public API (a user of the reference's ``src.osd.separation.Separator`` switches to this):

    from b200sep import Separator                      # same dataclass fields / separate() contract
    sep = Separator(backend="b200", device="cuda", checkpoint=path, n_src=3)
    w1, w2, w3 = sep.separate(chunk, sr)
    outs = sep.separate_batch(chunks, sr)              # all overlap segments in one forward
    res = pit_sisdr_batch(mixes, refs, preds)          # batched PIT SI-SDR / SI-SDRi
    OverlapBatchDriver(sep, eval_separation=True).run([MixtureJob(mix, sr, osd_segments, refs=refs)])
    b200sep.aux.resample_batch / frame_rms / build_gt_overlap_mask   # device versions of the code either side
"""
from . import aux, metrics, pipeline, sharding, synth, weights  # noqa: F401
from .pipeline import MixtureJob, OverlapBatchDriver  # noqa: F401
from ._native import LIB_PATH  # noqa: F401
from .pit import (PitResult, _compute_sdr_improvement_pit_k, _pit_best_si_sdr_k, _si_sdr,  # noqa: F401
                  compute_sdr_improvement_pit, compute_sdr_improvement_pit_k, pit_sisdr_batch,
                  pit_sisdr_device)
from .separator import BACKEND_NAME, SeparationEngine, Separator  # noqa: F401

__version__ = "0.1.0"
