"""Golden vectors of the featurisation step from the reference's OWN functions (build container only; needs /root/reference).

    python oracle/make_golden_featurize.py      # writes tests/golden/featurize.npz
Stores seeds + the reference's outputs; the polylines are regenerated from the seed (oracle/featurize_oracle.synth_wireframes).
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference")
from faceformer.datasets.data_para import sample_points  # noqa: E402  (the unmodified reference function)
from oracle.featurize_oracle import synth_wireframes  # noqa: E402

CASES = {"a": dict(n=5, num_lines=216, seed=1), "b": dict(n=3, num_lines=28, seed=2), "c": dict(n=2, num_lines=512, seed=3)}


def reference_featurize(wireframes, num_lines, P=50, D=2):
    """__getitem__ of ABCDataset_Parallel, input part (data_para.py:59-68), with the reference's sample_points."""
    inp = np.zeros((len(wireframes), num_lines, P, D), dtype=np.float32)
    mask = np.ones((len(wireframes), num_lines), dtype=bool)
    for w, edges in enumerate(wireframes):
        for i, edge in enumerate(edges):
            inp[w, i, :P] = sample_points(edge, P)
        mask[w, :len(edges)] = 0
    return inp, mask


out = {"meta": json.dumps({"cases": CASES, "numpy": np.__version__})}
for name, c in CASES.items():
    wfs = synth_wireframes(c["n"], c["num_lines"], c["seed"])
    inp, mask = reference_featurize(wfs, c["num_lines"])
    out[f"{name}_input"] = inp
    out[f"{name}_mask"] = mask
    print(name, inp.shape, int((~mask).sum()), "edges")
np.savez_compressed(os.path.join(ROOT, "tests", "golden", "featurize.npz"), **out)
