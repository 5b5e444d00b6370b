"""ORACLE helper (test infrastructure): run the reference's own Python files unchanged.

``install()`` registers ``oracle.asteroid_restated`` as ``asteroid`` / ``asteroid.models`` in
``sys.modules`` and stubs the other third-party engines the reference imports at module level
(``sherpa_onnx``, ``modelscope``), so that -- IN THE BUILD CONTAINER ONLY, where
``/root/reference`` exists -- ``from src.osd.separation import Separator`` and
``import overlap3_core`` / ``import evaluate_with_sources`` execute the reference's code
verbatim.  Used by ``oracle/make_golden.py`` and by tests that are skipped when
``/root/reference`` is absent (the GPU box).  Nothing here is imported by the product.
"""
from __future__ import annotations

import os
import sys
import types

REFERENCE_ROOT = "/root/reference"


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "src", "osd", "separation.py"))


def install(reference_root: str = REFERENCE_ROOT) -> None:
    from . import asteroid_restated as ar

    if "asteroid" not in sys.modules or not hasattr(sys.modules["asteroid"], "_b200_oracle_shim"):
        pkg = types.ModuleType("asteroid")
        pkg._b200_oracle_shim = True
        pkg.__path__ = []  # mark as package
        models = types.ModuleType("asteroid.models")
        models.ConvTasNet = ar.ConvTasNet
        models.BaseModel = ar.BaseModel
        pkg.models = models
        sys.modules["asteroid"] = pkg
        sys.modules["asteroid.models"] = models

    if "sherpa_onnx" not in sys.modules:
        sys.modules["sherpa_onnx"] = types.ModuleType("sherpa_onnx")
    if "modelscope" not in sys.modules:
        ms = types.ModuleType("modelscope")
        ms.__path__ = []
        msd = types.ModuleType("modelscope.msdatasets")

        class MsDataset:  # never used by the functions we call
            @staticmethod
            def load(*a, **k):
                raise RuntimeError("modelscope is stubbed in the oracle shim")

        msd.MsDataset = MsDataset
        ms.msdatasets = msd
        sys.modules["modelscope"] = ms
        sys.modules["modelscope.msdatasets"] = msd

    for p in (reference_root, os.path.join(reference_root, "scripts", "osd")):
        if os.path.isdir(p) and p not in sys.path:
            sys.path.insert(0, p)


def import_reference_separator():
    """The reference's own ``Separator`` class (src/osd/separation.py:14-163), unmodified."""
    install()
    import importlib.util
    spec = importlib.util.spec_from_file_location(
        "_ref_separation", os.path.join(REFERENCE_ROOT, "src", "osd", "separation.py"))
    mod = importlib.util.module_from_spec(spec)
    sys.modules["_ref_separation"] = mod
    spec.loader.exec_module(mod)
    return mod.Separator


def import_reference_pit():
    """(overlap3_core, evaluate_with_sources) reference modules, unmodified."""
    install()
    import importlib
    o3 = importlib.import_module("overlap3_core")
    ews = importlib.import_module("evaluate_with_sources")
    return o3, ews
