"""What the reference's Lightning harness does with `predict` right after the path.

  face_accuracy            Trainer.face_accuracy (faceformer/trainer.py:210-300): per-wireframe co-edge accuracy, duplicate removal,
                           face type by majority vote, precision / recall / type accuracy, token accuracy
  prediction_record /      the per-sample JSON test_step writes for the reconstruction scripts (trainer.py:118-136)
  write_prediction_json
  load_lightning_checkpoint  `Trainer.load_from_checkpoint` reduced to what the path needs (main.py:46, trainer.py:17-21): the
                           hyper-parameters' `model` node and the `model.`-prefixed state_dict
  load_yaml_config         configs/*.yml over the defaults of faceformer/config.py:7-52 (values only; no fvcore)

The per-SEQUENCE work (cut at the face-type token, offset removal, enclosedness walk, loop canonicalisation: trainer.py:196-206,
post_processing.py:8-20) runs on the GPU (ffb_parse_faces, one thread per sequence) for predictions AND labels; the per-WIREFRAME set
logic below touches a few dozen faces and stays on the host, like the reference.
"""
from __future__ import annotations

import json
import os
from collections import Counter
from typing import Dict, List, Sequence

import numpy as np

from .config import MODE_PARALLEL, ModelConfig


def map_coedge_into_edges(pairings: Dict[str, int], indices: Sequence[int]) -> List[int]:
    """post_processing.py:41-48: co-edge index -> edge index through the JSON's string-keyed pairing table."""
    return [pairings[str(i)] if str(i) in pairings else i for i in indices]


def _flatten(loops):
    return [i for loop in loops for i in loop]                  # utils.py:47-51


def parse_label_faces_host(labels: np.ndarray, token_len: int = 4, face_type_offset: int = 1):
    """Label half of Trainer.parse_parallel_faces (trainer.py:184-194): no `< num_edges` filter (used when is_coedge is off)."""
    faces = []
    for label in np.array(labels, dtype=np.int64):
        cut = np.where((label >= face_type_offset) & (label < token_len))[0] + 1
        label = np.split(label, cut)[0]
        face_type = label[-1] - face_type_offset
        label = label - token_len
        label = label[label >= 0]
        if len(label) > 0:
            faces.append((int(face_type), tuple(label.tolist())))
    return faces


def wireframe_metrics(predict_faces, label_faces, pairings=None, is_coedge: bool = True) -> dict:
    """The per-wireframe body of Trainer.face_accuracy (trainer.py:229-293) on already parsed (and, with is_coedge, enclosedness-filtered)
    faces: (face_type, loops) tuples when is_coedge else (face_type, indices)."""
    out = {}
    if is_coedge:
        face_tp = type_tp = 0
        label_set = set(label_faces)
        for pred_type, pred_face in predict_faces:                              # trainer.py:239-245
            for label_type, label_face in label_set:
                if pred_face == label_face:
                    face_tp += 1
                    if pred_type == label_type:
                        type_tp += 1
                    break
        if len(predict_faces) == 0:
            out["accuracy"], out["type_acc_coedge_seq"] = 0, 0
        else:
            out["accuracy"] = face_tp / len(predict_faces)
            out["type_acc_coedge_seq"] = 0 if face_tp == 0 else type_tp / face_tp
        pairings = pairings or {}
        label_faces = [(t, map_coedge_into_edges(pairings, _flatten(loops))) for t, loops in label_faces]      # trainer.py:257-258
        predict_faces = [(t, map_coedge_into_edges(pairings, _flatten(loops))) for t, loops in predict_faces]
    label_set = list(set((t, tuple(sorted(set(ind)))) for t, ind in label_faces))                              # trainer.py:261
    uniq: Dict[tuple, list] = {}
    for t, ind in predict_faces:                                                                               # trainer.py:264-270
        uniq.setdefault(tuple(sorted(set(ind))), []).append(t)
    predict_set = [(Counter(ts).most_common(1)[0][0], face) for face, ts in uniq.items()]                      # majority vote, trainer.py:272
    face_tp = type_tp = 0
    for pred_type, pred_face in predict_set:                                                                   # trainer.py:275-283
        for label_type, label_face in label_set:
            if pred_face == label_face:
                face_tp += 1
                if pred_type == label_type:
                    type_tp += 1
                break
    if len(predict_set) == 0 or len(label_set) == 0:
        out.update(precision=0, recall=0, type_acc=0)
    else:
        out.update(precision=face_tp / len(predict_set), recall=face_tp / len(label_set), type_acc=0 if face_tp == 0 else type_tp / face_tp)
    out.update(predictions=predict_set, labels=label_set)
    return out


def _subtract_like_the_parser(a: np.ndarray, token_len: int = 4, face_type_offset: int = 1) -> np.ndarray:
    """Trainer.parse_parallel_faces cuts every row after its first face-type token with np.split -- a VIEW -- and then does
    `row -= token.len` in place (trainer.py:190,202): by the time face_accuracy reaches its token accuracy (trainer.py:296-300) the cut
    prefix of every label / predict row has been shifted by -4 in the arrays themselves.  Reproduced here on a copy."""
    a = np.array(a, dtype=np.int64)
    flat = a.reshape(-1, a.shape[-1])
    for row in flat:
        hit = np.where((row >= face_type_offset) & (row < token_len))[0]
        end = int(hit[0]) + 1 if len(hit) else row.shape[0]
        row[:end] -= token_len
    return a


def token_accuracy(predicts: np.ndarray, labels: np.ndarray, pad: int = 0) -> float:
    """trainer.py:296-300, evaluated as the reference evaluates it: on the arrays its parser has already modified in place (see
    _subtract_like_the_parser).  `predicts` is [N,F,T] and `labels` [N,num_lines,T]: when F < num_lines the reference's
    `predicts == labels` is a shape-mismatched comparison, which the numpy it pins (1.19.5, environment.yml:25) evaluates to the scalar
    False -> 0 matches."""
    labels = _subtract_like_the_parser(labels)
    valid = labels > pad
    if predicts.shape != labels.shape:
        return 0.0
    predicts = _subtract_like_the_parser(predicts)
    return float((valid * (predicts == labels)).sum() / valid.sum())


def face_accuracy(engine, outputs: dict, raw_datas: Sequence[dict], is_coedge: bool = True, tol: float = 2e-4):
    """Trainer.face_accuracy (trainer.py:210-300).  outputs: the dict the model returned (`predict`, `label`, `id`); raw_datas[id] holds the
    dataset JSON's `edges` and `pairings`.  Returns (token accuracy, outputs) with the reference's list-valued keys added."""
    predicts, labels = outputs["predict"], outputs["label"]
    ids = [int(i) for i in (outputs["id"].tolist() if hasattr(outputs["id"], "tolist") else outputs["id"])]
    wfs = [raw_datas[i]["edges"] for i in ids]
    pred_faces = engine.parse_faces(predicts, wfs, tol=tol, check_enclosed=is_coedge)
    if is_coedge:           # labels through the same kernel: with the enclosedness walk an out-of-range label index is skipped either way
        label_faces = engine.parse_faces(labels, wfs, tol=tol, check_enclosed=True)
    else:
        lab = labels.cpu().numpy() if hasattr(labels, "cpu") else np.asarray(labels)
        label_faces = [parse_label_faces_host(lab[w]) for w in range(len(ids))]
    for k in ("precisions", "labels", "type_acc_coedge_seq", "recalls", "predictions", "accuracy", "type_acc"):
        outputs[k] = []
    for w, i in enumerate(ids):
        m = wireframe_metrics(pred_faces[w], label_faces[w], raw_datas[i].get("pairings"), is_coedge)
        if is_coedge:
            outputs["accuracy"].append(m["accuracy"])
            outputs["type_acc_coedge_seq"].append(m["type_acc_coedge_seq"])
        outputs["precisions"].append(m["precision"])
        outputs["recalls"].append(m["recall"])
        outputs["type_acc"].append(m["type_acc"])
        outputs["predictions"].append(m["predictions"])
        outputs["labels"].append(m["labels"])
    p = predicts.cpu().numpy() if hasattr(predicts, "cpu") else np.asarray(predicts)
    lab = labels.cpu().numpy() if hasattr(labels, "cpu") else np.asarray(labels)
    return token_accuracy(p, lab), outputs


# ---- wire formats (row f3) ------------------------------------------------------------------------------------------------
def prediction_record(raw_data: dict, predict_faces_w_types, label_faces_w_types) -> dict:
    """The dict test_step dumps per sample (trainer.py:127-132)."""
    return {"edges": raw_data["edges"], "dominant_directions": raw_data["dominant_directions"],
            "pred_faces": predict_faces_w_types, "label_faces": label_faces_w_types}


class _NumpyEncoder(json.JSONEncoder):          # what numpyencoder.NumpyEncoder does for the types that occur here
    def default(self, o):
        if isinstance(o, np.integer):
            return int(o)
        if isinstance(o, np.floating):
            return float(o)
        if isinstance(o, np.ndarray):
            return o.tolist()
        return super().default(o)


def write_prediction_json(log_dir: str, json_name: str, record: dict) -> str:
    """trainer.py:121,135-136: <log_dir>/json/<json_name[5:13]>.json"""
    os.makedirs(os.path.join(log_dir, "json"), exist_ok=True)
    path = os.path.join(log_dir, "json", f"{json_name[5:13]}.json")
    with open(path, "w") as f:
        json.dump(record, f, cls=_NumpyEncoder)
    return path


_MODEL_DEFAULTS = dict(num_points_per_line=50, num_lines=64, point_dim=2, label_seq_length=128, max_face_length=34, num_model=512, num_head=8,
                       num_feedforward=1024, num_encoder_layers=6, num_decoder_layers=6, dropout=0.2)      # config.py:27-39


def model_config_from_node(node: dict) -> ModelConfig:
    """cfg.model (config.py:27-49, or a checkpoint's hyper_parameters['model']) -> ModelConfig."""
    d = dict(_MODEL_DEFAULTS)
    d.update({k: v for k, v in dict(node).items() if k in d})
    tok = dict(node).get("token")
    ntok = int(tok["len"] if isinstance(tok, dict) else getattr(tok, "len", 4)) if tok is not None else 4
    return ModelConfig(num_token=ntok, **d)


def load_yaml_config(path: str) -> dict:
    """configs/*.yml over the defaults: returns dict(model=ModelConfig, model_class, dataset_class, post_process, raw=yaml dict)."""
    import yaml
    with open(path) as f:
        raw = yaml.safe_load(f) or {}
    pp = dict(enclosedness_tol=2e-4, is_coedge=True)                  # config.py:50-52
    pp.update(raw.get("post_process") or {})
    return dict(model=model_config_from_node(raw.get("model") or {}), model_class=raw.get("model_class", "SurfaceFormer"),
                dataset_class=raw.get("dataset_class", "ABCDataset"), post_process=pp, raw=raw)


def load_lightning_checkpoint(path: str):
    """A pytorch_lightning checkpoint as main.py:46 consumes it: (ModelConfig | None, mode | None, state_dict with the `model.` prefix
    stripped).  The hyper-parameters are saved by save_hyperparameters(cfg) (trainer.py:19) under `hyper_parameters`."""
    import torch
    ck = torch.load(path, map_location="cpu", weights_only=False)
    sd = {(k[6:] if k.startswith("model.") else k): v for k, v in ck["state_dict"].items()}
    hp = ck.get("hyper_parameters") or {}
    cfg = mode = None
    if "model" in hp:
        cfg = model_config_from_node(hp["model"])
        cls = hp.get("model_class", "")
        mode = MODE_PARALLEL if "Parallel" in str(cls) or "max_face_length" in dict(hp["model"]) and any(k.endswith("query_pos_enc.pos_embed.weight") and
                v.shape[0] == dict(hp["model"]).get("max_face_length") for k, v in sd.items()) else 1
    return cfg, mode, sd
