"""Golden vectors of the post-decode parsing step from the reference's OWN code (build container only; needs /root/reference).

Trainer.parse_parallel_faces cannot be imported (pytorch_lightning is absent), so its source text is cut out of
faceformer/trainer.py with `ast` and executed as-is with a stand-in `self`; filter_faces_by_encloseness is imported.

    python oracle/make_golden_faces.py      # writes tests/golden/faces.npz
"""
import ast
import json
import os
import sys
import textwrap
from types import SimpleNamespace

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference")
from faceformer.post_processing import filter_faces_by_encloseness  # noqa: E402  (unmodified reference function)
from oracle.faces_oracle import synth_case  # noqa: E402

src = open("/root/reference/faceformer/trainer.py").read()
fn = next(n for n in ast.walk(ast.parse(src)) if isinstance(n, ast.FunctionDef) and n.name == "parse_parallel_faces")
code = textwrap.dedent("\n".join(src.splitlines()[fn.lineno - 1:fn.end_lineno]))
ns = {"np": np}
exec(code, ns)
parse_parallel_faces = ns["parse_parallel_faces"]
SELF = SimpleNamespace(hparams=SimpleNamespace(model=SimpleNamespace(token=SimpleNamespace(PAD=0, SOS=1, SEP=2, EOS=3, len=4, face_type_offset=1))))

CASES = {"a": dict(n=6, num_lines=216, T=37, seed=1), "b": dict(n=9, num_lines=28, T=10, seed=2), "c": dict(n=4, num_lines=202, T=38, seed=3)}
TOL = 2e-4                                                   # config.py:51

out = {"meta": json.dumps({"cases": CASES, "tol": TOL})}
for name, c in CASES.items():
    wfs, pred = synth_case(c["n"], c["num_lines"], c["T"], c["seed"])
    parsed, filtered = [], []
    for w, edges in enumerate(wfs):
        labels = np.zeros((1, c["T"]), np.int64)             # the label half is not under test
        pf, _ = parse_parallel_faces(SELF, pred[w].copy(), labels, len(edges))
        pf = [(int(t), tuple(int(i) for i in idx)) for t, idx in pf]
        ff = filter_faces_by_encloseness(edges, pf, TOL)
        ff = [(int(t), tuple(tuple(int(i) for i in loop) for loop in loops)) for t, loops in ff]
        parsed.append(pf); filtered.append(ff)
    out[f"{name}_parsed"] = json.dumps(parsed)
    out[f"{name}_filtered"] = json.dumps(filtered)
    print(name, pred.shape, sum(len(p) for p in parsed), "parsed faces,", sum(len(f) for f in filtered), "enclosed")
np.savez_compressed(os.path.join(ROOT, "tests", "golden", "faces.npz"), **out)
