"""Overlay of the reference's ``faceformer.models`` package (drop-in mechanism).

/root/reference/faceformer/ has no ``__init__.py``: it is a namespace package whose portions
merge across ``sys.path``, while ``faceformer/models/`` is a regular package, so the FIRST one
on the path wins.  With this repo ahead of the reference on ``sys.path`` (run a copy of the
reference's ``main.py`` from this repo root with ``PYTHONPATH=<reference>``), ``main.py``'s
``from faceformer.models import *`` + ``str_to_class(cfg.model_class)`` (main.py:9,13-14,29)
resolves

    model_class: 'SurfaceFormer_Parallel_B200'   /   'SurfaceFormer_B200'

to the B200 classes, while ``faceformer.trainer``, ``faceformer.config``, ``faceformer.datasets``
and ``faceformer.embedding`` still come from the reference.  If the reference is importable its
own two model classes are re-exported unchanged, so every existing config keeps working.
"""
import os as _os
import sys as _sys

from faceformer_b200.models import SurfaceFormer_B200, SurfaceFormer_Parallel_B200  # noqa: F401

__all__ = ["SurfaceFormer_B200", "SurfaceFormer_Parallel_B200"]

for _p in list(_sys.path) + [_os.environ.get("FACEFORMER_REFERENCE", "")]:
    _cand = _os.path.join(_p, "faceformer", "models") if _p else ""
    if _cand and _os.path.isfile(_os.path.join(_cand, "model_para.py")) and \
            _os.path.abspath(_cand) != _os.path.dirname(_os.path.abspath(__file__)):
        __path__.append(_cand)          # lets `.model` / `.model_para` resolve to the reference's files
        try:
            from .model import SurfaceFormer  # noqa: F401
            from .model_para import SurfaceFormer_Parallel  # noqa: F401
            __all__ += ["SurfaceFormer", "SurfaceFormer_Parallel"]
        except Exception:  # reference present but not importable (missing deps): B200 classes only
            pass
        break
