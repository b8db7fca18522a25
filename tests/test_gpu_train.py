"""Row f4: the teacher-forced forward pass (ffb_forward_train) against outputs of the unmodified reference's forward_train
(tests/golden/train_forward_*.npz, oracle/make_golden_train.py): pointer within the logit tolerance, and the quantities Trainer.compute_loss
derives from it (trainer.py:61-79) -- loss, token accuracy, argmax predictions."""
import json
import os

import numpy as np
import pytest
import torch

from faceformer_b200 import synth
from faceformer_b200.config import MODE_PARALLEL, MODE_SEQ2SEQ, ModelConfig
from faceformer_b200.engine import Engine
from faceformer_b200.lib import FFBError
from oracle import faceformer_oracle as orc
from util import GOLDEN, LOGIT_TOL, logits_close

pytestmark = pytest.mark.gpu


def load(name):
    with np.load(os.path.join(GOLDEN, name + ".npz")) as z:
        g = {k: z[k] for k in z.files}
    meta = json.loads(str(g.pop("meta")))
    cfg = ModelConfig(**meta["cfg"])
    mode = meta.get("mode", MODE_PARALLEL)
    if mode == MODE_PARALLEL:
        sd = synth.load_state_dict_npz(os.path.join(GOLDEN, meta["weights"]))
        batch = synth.polygon_batch(cfg, meta["n"], seed=meta["seed"])
    else:
        sd = synth.synth_state_dict(cfg, mode, meta["weights"][1], "diverse")
        batch = synth.seq2seq_labels(cfg, synth.synth_batch(cfg, mode, meta["n"], meta["seed"], lo=5, hi=min(60, cfg.num_lines)), meta["seed"])
    g.update(cfg=cfg, mode=mode, meta=meta, sd=sd, batch=batch)
    return g


@pytest.mark.parametrize("name", ["train_forward_tiny", "train_forward_mid", "train_forward_seq2seq_tiny", "train_forward_seq2seq"])
@pytest.mark.parametrize("device_path", [True, False])
def test_teacher_forced_pointer_loss_and_accuracy_match_the_reference(name, device_path):
    g = load(name)
    b, mode = g["batch"], g["mode"]
    e = Engine(g["cfg"], mode, 0)
    e.load_state_dict(g["sd"])
    coords = b["input"].reshape(b["input"].shape[0], b["input"].shape[1], -1)
    args = [coords, b["input_mask"], b["num_input"] if mode == MODE_PARALLEL else None, b["label"], b["label_mask"]]
    if device_path:
        args = [None if a is None else torch.from_numpy(np.ascontiguousarray(a)).cuda() for a in args]
    ptr, mem = e.forward_train(*args, want_embedding=True)
    if device_path:
        ptr, mem = ptr.cpu().numpy(), mem.cpu().numpy()
    F = int(b["num_input"].max()) if mode == MODE_PARALLEL else 1
    T = g["cfg"].seq_len(mode)
    assert ptr.shape == (len(b["num_input"]) * F, T - 1, g["cfg"].num_model)
    st, ps = g["meta"]["seq_step"], g["meta"].get("pos_step", 1)
    ok, d = logits_close(ptr[::st, ::ps], g["pointer"], b64=g["pointer64"])
    assert ok, f"pointer differs from the reference's by {d}"
    assert mem.shape == g["memory"].shape                          # outputs['embedding'] per wireframe, rows of padded edges included
    assert np.max(np.abs(mem - g["memory"])) <= LOGIT_TOL
    # what the trainer computes from it (trainer.py:61-79)
    out = dict(embedding=np.repeat(mem, F, axis=0), pointer=ptr, label=g["label"])
    loss, acc, pred = orc.teacher_forced_loss(out)
    # mean cross-entropy over ~10^3 tokens of logits = <memory row, pointer> (512 terms each): both sides are fp32-class evaluations of the
    # pointer (ours: fp16x2 tensor-core GEMMs, float64 encoder), so the loss agrees to ~1e-4 relative, not to fp32 round-off
    assert abs(loss - float(g["loss"])) <= 1e-4 * max(1.0, abs(float(g["loss"])))
    valid = g["label"] != 0
    assert np.array_equal(pred[valid], g["pred"][valid])
    assert abs(acc - float(g["acc"])) < 1e-9
    e.close()


def test_teacher_forced_pass_rejects_tokens_outside_the_memory():
    g = load("train_forward_tiny")
    b = g["batch"]
    e = Engine(g["cfg"], MODE_PARALLEL, 0)
    e.load_state_dict(g["sd"])
    bad = b["label"].copy()
    bad[0, 0, 1] = g["cfg"].mem_len + 3
    with pytest.raises(FFBError):
        e.forward_train(b["input"].reshape(b["input"].shape[0], b["input"].shape[1], -1), b["input_mask"], b["num_input"], bad, b["label_mask"])
    # the handle stays usable: a regular decode afterwards
    pred, steps = e.forward_eval(b["input"].reshape(b["input"].shape[0], b["input"].shape[1], -1), b["input_mask"], b["num_input"])
    assert steps >= 1 and pred.shape[0] == len(b["num_input"])
    e.close()


def test_model_class_forward_train_feeds_compute_loss():
    """SurfaceFormer_Parallel_B200 in train() mode: forward(dict) returns the reference's training outputs; the loss of
    Trainer.compute_loss (trainer.py:61-79), evaluated with torch on the GPU, equals the reference's."""
    from faceformer_b200.models import SurfaceFormer_Parallel_B200
    g = load("train_forward_mid")
    m = SurfaceFormer_Parallel_B200(**g["cfg"].model_kwargs(MODE_PARALLEL)).cuda()
    m.load_state_dict({k: torch.from_numpy(np.asarray(v)) for k, v in g["sd"].items()}, strict=True)
    m.train()
    out = m({k: torch.from_numpy(v).cuda() for k, v in g["batch"].items()})
    logits = torch.bmm(out["embedding"], out["pointer"].transpose(1, 2))
    labels = out["label"]
    assert np.array_equal(labels.cpu().numpy(), g["label"])
    loss = torch.nn.functional.cross_entropy(logits, labels, ignore_index=0, reduction="sum") / (labels != 0).sum()
    assert abs(float(loss) - float(g["loss"])) <= 1e-4 * max(1.0, float(g["loss"]))
    pred = torch.argmax(logits, dim=1).cpu().numpy()
    valid = g["label"] != 0
    assert np.array_equal(pred[valid], g["pred"][valid])
