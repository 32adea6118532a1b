"""Import the UNMODIFIED reference from /root/reference (build container only).  TEST INFRASTRUCTURE.

``models.py`` imports as-is.  ``metrics.py`` / ``train_eval.py`` import packages that are not
installed in this image (pycocotools, skimage, matplotlib, seaborn); empty stub modules are
registered for those names only - none of the hot-path functions touch them.
``/root/reference`` does not exist on the GPU box: callers must check ``available()``.
"""
from __future__ import annotations

import contextlib
import io
import os
import sys
import types

REF_DIR = "/root/reference"


def available() -> bool:
    return os.path.isfile(os.path.join(REF_DIR, "models.py"))


def _stub(name: str):
    if name in sys.modules:
        return sys.modules[name]
    m = types.ModuleType(name)
    m.__path__ = []  # behave like a package
    sys.modules[name] = m
    return m


def load():
    """Returns (models, metrics, train_eval) reference modules."""
    if not available():
        raise RuntimeError("reference not present at /root/reference")
    for n in ("pycocotools", "pycocotools.mask", "pycocotools.coco", "pycocotools.cocoeval",
              "skimage", "skimage.measure", "skimage.feature",
              "matplotlib", "matplotlib.pyplot", "matplotlib.font_manager", "matplotlib.patches",
              "matplotlib.colors", "matplotlib.gridspec", "seaborn"):
        try:
            __import__(n)
        except Exception:
            _stub(n)
    sys.modules["pycocotools.coco"].__dict__.setdefault("COCO", object)
    sys.modules["pycocotools.cocoeval"].__dict__.setdefault("COCOeval", object)
    sys.modules["skimage.feature"].__dict__.setdefault("peak_local_max", None)
    sys.modules["skimage"].__dict__.setdefault("measure", sys.modules["skimage.measure"])
    if REF_DIR not in sys.path:
        sys.path.insert(0, REF_DIR)
    saved = {k: sys.modules.get(k) for k in ("models", "metrics", "train_eval", "dataset", "visualization")}
    for k in saved:
        sys.modules.pop(k, None)
    try:
        with contextlib.redirect_stdout(io.StringIO()):
            import models as ref_models  # type: ignore
            import metrics as ref_metrics  # type: ignore
            import train_eval as ref_train_eval  # type: ignore
    finally:
        # do not leave the reference's module names shadowing anything else
        for k, v in saved.items():
            if v is not None:
                sys.modules[k] = v
            else:
                sys.modules.pop(k, None)
        if REF_DIR in sys.path:
            sys.path.remove(REF_DIR)
    return ref_models, ref_metrics, ref_train_eval
