"""Evaluation bookkeeping that downstream consumers of the reference rely on.

* ``agg_stats``      -- ``_agg`` of ``/root/reference/scripts/osd/overlap3_core.py:860-869``
                        (float32 array, mean / median / population std, rounded to 4 dp)
* ``sep_metrics``    -- the ``sep_*`` keys of ``metrics.json`` (``overlap3_core.py:911-925``), which
                        ``/root/reference/batch_eval.py:65-84`` averages across runs
* ``sep_detail_row`` -- one ``overlap_sep_details.csv`` row (``overlap3_core.py:732-742``)
"""
from __future__ import annotations

from typing import Any, Dict, List, Optional, Sequence

import numpy as np


def agg_stats(vals: Sequence[float]) -> Dict[str, Optional[float]]:
    if not len(vals):
        return {"mean": None, "median": None, "std": None, "count": 0}
    arr = np.asarray(vals, dtype=np.float32)
    return {"mean": round(float(np.mean(arr)), 4), "median": round(float(np.median(arr)), 4),
            "std": round(float(np.std(arr)), 4), "count": int(arr.size)}


def sep_metrics(sisdr: Sequence[float], sisdri: Sequence[float]) -> Dict[str, Any]:
    """NaN rows are dropped first, as the call site does (``overlap3_core.py:729``)."""
    keep = [i for i, (a, b) in enumerate(zip(sisdr, sisdri)) if not (np.isnan(a) or np.isnan(b))]
    s = agg_stats([float(sisdr[i]) for i in keep])
    si = agg_stats([float(sisdri[i]) for i in keep])
    return {"sep_eval_k_refs": None, "sep_eval_segments": s["count"],
            "sep_sisdr_mean": s["mean"], "sep_sisdr_median": s["median"], "sep_sisdr_std": s["std"],
            "sep_sisdri_mean": si["mean"], "sep_sisdri_median": si["median"], "sep_sisdri_std": si["std"]}


def sep_detail_row(mix_path: str, s: float, e: float, K: int, best: float, sdri: float,
                   idx_sel: Sequence[int]) -> List[Any]:
    return [mix_path, f"{s:.3f}", f"{e:.3f}", K, f"{best:.4f}", f"{sdri:.4f}",
            ";".join(str(int(i)) for i in idx_sel)]
