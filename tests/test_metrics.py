"""SURVEY.md 8 rows f2 (set-level metrics of Trainer.face_accuracy, trainer.py:210-300) and f3 (wire formats: per-sample JSON,
YAML config values, Lightning checkpoint).  Expected values come from the reference's own code (oracle/make_golden_metrics.py)."""
import json
import os

import numpy as np
import pytest

from faceformer_b200 import metrics
from oracle import faces_oracle as fo
from oracle.metrics_oracle import synth_metrics_case

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "metrics.npz")


def load():
    with np.load(GOLDEN) as z:
        meta = json.loads(str(z["meta"]))
        return meta, {k: json.loads(str(z[k])) for k in z.files if k != "meta"}


def _check(got_list, want, is_coedge):
    for key_g, key_w in (("precision", "precisions"), ("recall", "recalls"), ("type_acc", "type_acc")):
        assert [float(m[key_g]) for m in got_list] == pytest.approx(want[key_w], abs=0)
    if is_coedge:
        assert [float(m["accuracy"]) for m in got_list] == pytest.approx(want["accuracy"], abs=0)
        assert [float(m["type_acc_coedge_seq"]) for m in got_list] == pytest.approx(want["type_acc_coedge_seq"], abs=0)
    for m, wp, wl in zip(got_list, want["predictions"], want["labels"]):
        assert [[int(t), list(f)] for t, f in m["predictions"]] == wp                       # majority-voted, de-duplicated, insertion order
        assert sorted([int(t), list(f)] for t, f in m["labels"]) == wl


@pytest.mark.parametrize("name", ["a", "b"])
@pytest.mark.parametrize("is_coedge", [True, False])
def test_set_level_metrics_match_reference_code(name, is_coedge):
    """Host half (duplicate removal, majority vote, precision / recall, co-edge mapping) on faces parsed by the CPU oracle."""
    meta, gold = load()
    c, tol = meta["cases"][name], meta["tol"]
    raw, pred, lab = synth_metrics_case(c["n"], c["num_lines"], c["T"], c["seed"])
    got = []
    for w, rd in enumerate(raw):
        edges = rd["edges"]
        pf = fo.parse_predicts(pred[w], len(edges))
        if is_coedge:
            pf = fo.filter_faces_by_encloseness(edges, pf, tol)
            lf = fo.filter_faces_by_encloseness(edges, metrics.parse_label_faces_host(lab[w]), tol)
        else:
            lf = metrics.parse_label_faces_host(lab[w])
        got.append(metrics.wireframe_metrics(pf, lf, rd["pairings"], is_coedge))
    want = gold[f"{name}_{int(is_coedge)}"]
    _check(got, want, is_coedge)
    assert metrics.token_accuracy(pred, lab) == pytest.approx(want["token_acc"], abs=1e-15)
    assert metrics.token_accuracy(pred[:, :-1], lab) == 0.0          # F < num_lines: the reference's mismatched comparison counts nothing


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["a", "b"])
@pytest.mark.parametrize("is_coedge", [True, False])
@pytest.mark.parametrize("device", [True, False])
def test_face_accuracy_on_the_gpu_matches_reference_code(name, is_coedge, device):
    """The whole face_accuracy: per-sequence parsing + enclosedness on the GPU (predictions and labels), set logic on the host."""
    import torch
    from faceformer_b200.config import MODE_PARALLEL, OURS
    from faceformer_b200.engine import Engine
    meta, gold = load()
    c, tol = meta["cases"][name], meta["tol"]
    raw, pred, lab = synth_metrics_case(c["n"], c["num_lines"], c["T"], c["seed"])
    e = Engine(OURS.replace(num_lines=c["num_lines"], max_face_length=c["T"]), MODE_PARALLEL, 0)
    outputs = {"predict": torch.from_numpy(pred).cuda() if device else pred, "label": torch.from_numpy(lab).cuda() if device else lab,
               "id": list(range(len(raw)))}
    acc, out = metrics.face_accuracy(e, outputs, raw, is_coedge=is_coedge, tol=tol)
    want = gold[f"{name}_{int(is_coedge)}"]
    got = [dict(precision=out["precisions"][w], recall=out["recalls"][w], type_acc=out["type_acc"][w],
                accuracy=out["accuracy"][w] if is_coedge else 0, type_acc_coedge_seq=out["type_acc_coedge_seq"][w] if is_coedge else 0,
                predictions=out["predictions"][w], labels=out["labels"][w]) for w in range(len(raw))]
    _check(got, want, is_coedge)
    assert acc == pytest.approx(want["token_acc"], abs=1e-15)
    e.close()


def test_prediction_json_and_config_and_checkpoint(tmp_path):
    import torch
    import yaml
    from faceformer_b200 import synth
    from faceformer_b200.config import MODE_PARALLEL, TINY
    from faceformer_b200.engine import pack_state_dict
    # per-sample JSON (trainer.py:118-136)
    raw = {"edges": [[[0.0, 0.0], [1.0, 0.0]]], "dominant_directions": [[1, 0, 0]], "faces": []}
    rec = metrics.prediction_record(raw, [(np.int64(0), (0, 1, 2))], [(0, (0, 1, 2))])
    path = metrics.write_prediction_json(str(tmp_path), "json/00001234_abc.json", rec)
    assert os.path.basename(path) == "00001234.json"
    back = json.load(open(path))
    assert back["pred_faces"] == [[0, [0, 1, 2]]] and back["edges"] == raw["edges"] and set(back) == {"edges", "dominant_directions", "pred_faces", "label_faces"}
    # YAML overrides over config.py defaults (configs/ours.yml:19-26)
    y = tmp_path / "ours.yml"
    y.write_text(yaml.safe_dump({"model_class": "SurfaceFormer_Parallel", "dataset_class": "ABCDataset_Parallel",
                                 "model": {"num_lines": 216, "max_face_length": 37}, "post_process": {"is_coedge": True}}))
    cfg = metrics.load_yaml_config(str(y))
    assert cfg["model"].num_lines == 216 and cfg["model"].max_face_length == 37 and cfg["model"].num_model == 512
    assert cfg["model_class"] == "SurfaceFormer_Parallel" and cfg["post_process"]["enclosedness_tol"] == 2e-4
    # Lightning checkpoint: `model.`-prefixed state_dict + hyper_parameters (main.py:46, trainer.py:17-21)
    sd = synth.synth_state_dict(TINY, MODE_PARALLEL, 0, "diverse")
    ck = {"state_dict": {"model." + k: torch.from_numpy(np.asarray(v)) for k, v in sd.items()},
          "hyper_parameters": {"model_class": "SurfaceFormer_Parallel", "model": dict(TINY.model_kwargs(MODE_PARALLEL), token={"len": 4})}}
    p = tmp_path / "last.ckpt"
    torch.save(ck, str(p))
    cfg2, mode, sd2 = metrics.load_lightning_checkpoint(str(p))
    assert cfg2 == TINY.replace(label_seq_length=cfg2.label_seq_length) and mode == MODE_PARALLEL
    assert np.array_equal(pack_state_dict(sd2, TINY, MODE_PARALLEL), pack_state_dict(sd, TINY, MODE_PARALLEL))
