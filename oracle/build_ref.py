"""Recipe for oracle/_ref/: the UNMODIFIED reference model code, importable on the GPU box.

TEST INFRASTRUCTURE.  The reference is pure Python (no setup.py, nothing to compile or pip-install), and
/root/reference does not exist on the GPU box.  This script copies -- byte for byte, verified by sha256 -- exactly the
files the hot path needs from where they lie under /root/reference into oracle/_ref/ (git-ignored:
reference sources never enter this repo's history; not gpurun-ignored: the directory travels to the GPU box like a built
.so).  Only `bench.py --impl reference` / the `cpu_baseline` leg (timing the reference's own CPU implementation,
kind "reference") and tests/ (as the checker) import it.

    python oracle/build_ref.py            # -> oracle/_ref/faceformer/{transformer,embedding,utils}.py, models/*.py, MANIFEST.json
"""
from __future__ import annotations

import hashlib
import json
import os
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("FACEFORMER_REFERENCE", "/root/reference")
OUT = os.path.join(ROOT, "oracle", "_ref")
FILES = ["faceformer/transformer.py", "faceformer/embedding.py", "faceformer/utils.py",
         "faceformer/models/__init__.py", "faceformer/models/model.py", "faceformer/models/model_para.py"]


def sha(path):
    with open(path, "rb") as f:
        return hashlib.sha256(f.read()).hexdigest()


def build(verbose: bool = True) -> bool:
    """Returns True if oracle/_ref is complete afterwards.  Without /root/reference an existing copy is kept as it is."""
    if not os.path.isdir(REF):
        ok = all(os.path.isfile(os.path.join(OUT, f)) for f in FILES)
        if verbose:
            print(f"oracle/_ref: {REF} not present; existing copy {'complete' if ok else 'MISSING'}")
        return ok
    manifest = {}
    for rel in FILES:
        src, dst = os.path.join(REF, rel), os.path.join(OUT, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(src, dst)
        assert sha(src) == sha(dst)
        manifest[rel] = sha(dst)
    with open(os.path.join(OUT, "MANIFEST.json"), "w") as f:
        json.dump({"source": REF, "sha256": manifest}, f, indent=1)
    if verbose:
        print(f"oracle/_ref: copied {len(FILES)} unmodified reference files from {REF}")
    return True


def import_reference():
    """-> (SurfaceFormer, SurfaceFormer_Parallel) of the unmodified reference in oracle/_ref (it must precede this repo's
    `faceformer/models` overlay on sys.path: `faceformer` is a namespace package, `faceformer.models` a regular one)."""
    if not all(os.path.isfile(os.path.join(OUT, f)) for f in FILES):
        raise RuntimeError("oracle/_ref is missing: run `python oracle/build_ref.py` in the build container")
    for name in [m for m in sys.modules if m == "faceformer" or m.startswith("faceformer.")]:
        del sys.modules[name]
    sys.path.insert(0, OUT)
    try:
        from faceformer.models import SurfaceFormer, SurfaceFormer_Parallel
    finally:
        sys.path.remove(OUT)
    assert os.path.abspath(sys.modules["faceformer.models"].__file__).startswith(OUT)
    return SurfaceFormer, SurfaceFormer_Parallel


if __name__ == "__main__":
    sys.exit(0 if build() else 1)
