"""Golden vectors of the set-level metrics from the reference's OWN code (build container only; needs /root/reference).

Trainer cannot be imported (pytorch_lightning is absent), so face_accuracy, parse_parallel_faces and parse_faces are cut out of
faceformer/trainer.py with `ast` and executed as they are with a stand-in `self`; filter_faces_by_encloseness,
map_coedge_into_edges and flatten_list are imported from the reference.  F == num_lines in every case so that the reference's final
`predicts == labels` comparison (trainer.py:297) is well-formed under numpy >= 1.25 (SURVEY.md 8f2 "trap").

    python oracle/make_golden_metrics.py      # writes tests/golden/metrics.npz
"""
import ast
import json
import os
import sys
import textwrap
from collections import Counter
from types import SimpleNamespace

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference")
from faceformer.post_processing import filter_faces_by_encloseness, map_coedge_into_edges  # noqa: E402  (unmodified reference functions)
from faceformer.utils import flatten_list  # noqa: E402
from oracle.metrics_oracle import synth_metrics_case  # noqa: E402

src = open("/root/reference/faceformer/trainer.py").read()
ns = {"np": np, "Counter": Counter, "filter_faces_by_encloseness": filter_faces_by_encloseness,
      "map_coedge_into_edges": map_coedge_into_edges, "flatten_list": flatten_list}
for fname in ("face_accuracy", "parse_parallel_faces", "parse_faces"):
    fn = next(n for n in ast.walk(ast.parse(src)) if isinstance(n, ast.FunctionDef) and n.name == fname)
    exec(textwrap.dedent("\n".join(src.splitlines()[fn.lineno - 1:fn.end_lineno])), ns)

CASES = {"a": dict(n=5, num_lines=28, T=10, seed=11), "b": dict(n=3, num_lines=216, T=37, seed=12)}
TOL = 2e-4


def run(raw, pred, lab, is_coedge):
    token = SimpleNamespace(PAD=0, SOS=1, SEP=2, EOS=3, len=4, face_type_offset=1)
    me = SimpleNamespace(hparams=SimpleNamespace(model=SimpleNamespace(token=token),
                                                 post_process=SimpleNamespace(is_coedge=is_coedge, enclosedness_tol=TOL)),
                         dataset=SimpleNamespace(raw_datas=raw))
    me.parse_parallel_faces = lambda p, l, ne: ns["parse_parallel_faces"](me, p, l, ne)
    me.parse_faces = lambda p, l, ne: ns["parse_faces"](me, p, l, ne)
    outputs = {"label": torch.from_numpy(lab.copy()), "predict": torch.from_numpy(pred.copy()), "id": list(range(len(raw)))}
    acc, out = ns["face_accuracy"](me, outputs)
    clean = lambda faces: [[int(t), [int(i) for i in idx]] for t, idx in faces]
    return {"token_acc": float(acc), "precisions": [float(x) for x in out["precisions"]], "recalls": [float(x) for x in out["recalls"]],
            "type_acc": [float(x) for x in out["type_acc"]], "accuracy": [float(x) for x in out["accuracy"]],
            "type_acc_coedge_seq": [float(x) for x in out["type_acc_coedge_seq"]],
            "predictions": [clean(p) for p in out["predictions"]], "labels": [sorted(clean(p)) for p in out["labels"]]}


out = {"meta": json.dumps({"cases": CASES, "tol": TOL})}
for name, c in CASES.items():
    raw, pred, lab = synth_metrics_case(c["n"], c["num_lines"], c["T"], c["seed"])
    F = pred.shape[1]
    for is_coedge in (True, False):
        r = run(raw, pred, lab, is_coedge)
        out[f"{name}_{int(is_coedge)}"] = json.dumps(r)
        print(name, "is_coedge", is_coedge, "F", F, "token_acc %.4f" % r["token_acc"], "precision", np.round(r["precisions"], 3), "recall", np.round(r["recalls"], 3))
np.savez_compressed(os.path.join(ROOT, "tests", "golden", "metrics.npz"), **out)
