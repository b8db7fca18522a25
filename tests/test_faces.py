"""prediction parsing + enclosedness filter (trainer.py:196-206, post_processing.py:8-20,
check_faces_enclosed.py:11-46).  The oracle restatement is pinned to outputs of the reference's own code
(tests/golden/faces.npz); the CUDA kernel must reproduce them exactly (integer work: identical faces, loops, order)."""
import json
import os

import numpy as np
import pytest

from oracle import faces_oracle as fo

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "faces.npz")


def _tup(x):
    return tuple(_tup(i) for i in x) if isinstance(x, list) else x


def cases():
    with np.load(GOLDEN) as z:
        meta = json.loads(str(z["meta"]))
        return meta["tol"], {k: (c, [[_tup(f) for f in w] for w in json.loads(str(z[f"{k}_parsed"]))],
                                 [[_tup(f) for f in w] for w in json.loads(str(z[f"{k}_filtered"]))]) for k, c in meta["cases"].items()}


@pytest.mark.parametrize("name", ["a", "b", "c"])
def test_oracle_matches_reference_code(name):
    tol, cs = cases()
    c, parsed, filtered = cs[name]
    wfs, pred = fo.synth_case(c["n"], c["num_lines"], c["T"], c["seed"])
    for w, edges in enumerate(wfs):
        pf = fo.parse_predicts(pred[w], len(edges))
        assert pf == parsed[w]
        assert fo.filter_faces_by_encloseness(edges, pf, tol) == filtered[w]
    assert sum(len(f) for f in filtered) > 0 and sum(len(f) for f in filtered) < sum(len(p) for p in parsed)


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["a", "b", "c"])
@pytest.mark.parametrize("device", [True, False])
def test_cuda_parse_faces_matches_reference_code(name, device):
    import torch
    from faceformer_b200.config import MODE_PARALLEL, OURS
    from faceformer_b200.engine import Engine
    tol, cs = cases()
    c, parsed, filtered = cs[name]
    wfs, pred = fo.synth_case(c["n"], c["num_lines"], c["T"], c["seed"])
    e = Engine(OURS.replace(num_lines=c["num_lines"], max_face_length=c["T"]), MODE_PARALLEL, 0)
    p = torch.from_numpy(pred).cuda() if device else pred
    assert e.parse_faces(p, wfs, tol=tol, check_enclosed=False) == parsed
    assert e.parse_faces(p, wfs, tol=tol, check_enclosed=True) == filtered
    e.close()


@pytest.mark.gpu
def test_parse_faces_after_a_real_decode():
    """decode the trained tiny checkpoint on polygon wireframes, parse on the device: equals the oracle on the same predict."""
    import torch
    from faceformer_b200.engine import Engine
    from util import load_case
    g = load_case("tiny_parallel_trained_b")
    b = g["batch"]
    e = Engine(g["cfg"], g["mode"], 0)
    e.load_state_dict(g["sd"])
    pred, _ = e.forward_eval(torch.from_numpy(b["input"]).cuda().flatten(2), torch.from_numpy(b["input_mask"]).cuda(),
                             torch.from_numpy(b["num_input"]).cuda())
    wfs = [[[pt.tolist() for pt in b["input"][w, i].astype(np.float64)] for i in range(int(b["num_input"][w]))] for w in range(len(b["num_input"]))]
    got = e.parse_faces(pred, wfs, tol=2e-4)
    pn = pred.cpu().numpy()
    want = [fo.filter_faces_by_encloseness(wfs[w], fo.parse_predicts(pn[w], len(wfs[w])), 2e-4) for w in range(len(wfs))]
    assert got == want and sum(len(f) for f in got) > 0
    e.close()


@pytest.mark.gpu
def test_model_class_featurize_decode_parse_roundtrip():
    """The reference-facing class: raw edge lists -> featurize -> model(batch) -> parse_faces, all on the device; equals the
    oracle's featurisation / parsing around the same decode."""
    import torch
    from faceformer_b200 import synth
    from faceformer_b200.config import MODE_PARALLEL, TINY
    from faceformer_b200.models import SurfaceFormer_Parallel_B200
    from oracle import featurize_oracle as feo
    wfs, _ = fo.synth_case(5, TINY.num_lines, TINY.max_face_length, 21)
    sd = synth.synth_state_dict(TINY, MODE_PARALLEL, 3, "diverse")
    m = SurfaceFormer_Parallel_B200(**TINY.model_kwargs(MODE_PARALLEL)).eval()
    m.load_state_dict({k: torch.from_numpy(np.asarray(v)) for k, v in sd.items()}, strict=True)
    m = m.cuda()
    inp, mask, ni = m.featurize(wfs)
    oi, om, on = feo.featurize(wfs, TINY.num_lines)
    assert np.array_equal(inp.cpu().numpy().view(np.uint32), oi.view(np.uint32)) and np.array_equal(mask.cpu().numpy(), om)
    batch = {"input": inp, "input_mask": mask, "num_input": ni,
             "label": torch.zeros((len(wfs), TINY.num_lines, TINY.max_face_length), dtype=torch.int64, device="cuda")}
    with torch.no_grad():
        out = m(batch)
    got = m.parse_faces(out["predict"], wfs)
    pn = out["predict"].cpu().numpy()
    want = [fo.filter_faces_by_encloseness(wfs[w], fo.parse_predicts(pn[w], len(wfs[w])), 2e-4) for w in range(len(wfs))]
    assert got == want
