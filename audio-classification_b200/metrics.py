"""Evaluation bookkeeping that downstream consumers of the reference rely on.

* ``agg_stats``      -- ``_agg`` of ``/root/reference/scripts/osd/overlap3_core.py:860-869``
                        (float32 array, mean / median / population std, rounded to 4 dp)
* ``sep_metrics``    -- the ``sep_*`` keys of ``metrics.json`` (``overlap3_core.py:911-925``), which
                        ``/root/reference/batch_eval.py:65-84`` averages across runs
* ``sep_detail_row`` -- one ``overlap_sep_details.csv`` row (``overlap3_core.py:732-742``)
* ``write_sep_details_csv`` -- the file itself, header included
                        (``/root/reference/scripts/osd/offline_overlap_3src.py:209-228``)
* ``merge_metrics``  -- ``metrics.update(...)`` of the ``sep_*`` keys into a metrics.json dict
                        (``overlap3_core.py:911-925``), written with the reference's json.dump settings
"""
from __future__ import annotations

import csv
import json
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


SEP_DETAILS_HEADER = ["wav", "start", "end", "k_refs", "sisdr", "sisdri", "selected_pred_indices"]


def write_sep_details_csv(path: str, rows: Sequence[Sequence[Any]]) -> None:
    """offline_overlap_3src.py:209-228: utf-8, ``newline=""``, header row, then ``sep_details_rows`` verbatim."""
    with open(path, "w", newline="", encoding="utf-8") as fh:
        w = csv.writer(fh)
        w.writerow(SEP_DETAILS_HEADER)
        for row in rows:
            w.writerow(row)


def merge_metrics(metrics: Dict[str, Any], sisdr: Sequence[float], sisdri: Sequence[float],
                  time_sep_sec: Optional[float] = None) -> Dict[str, Any]:
    """Adds the ``sep_*`` keys (and ``time_sep_sec`` rounded to 3 dp, ``overlap3_core.py:903``) to a metrics dict."""
    out = dict(metrics)
    if time_sep_sec is not None:
        out["time_sep_sec"] = round(float(time_sep_sec), 3)
    out.update(sep_metrics(sisdr, sisdri))
    return out


def write_metrics_json(path: str, metrics: Dict[str, Any]) -> None:
    """offline_overlap_3src.py writes metrics.json with ``ensure_ascii=False, indent=2``."""
    with open(path, "w", encoding="utf-8") as f:
        json.dump(metrics, f, ensure_ascii=False, indent=2)
