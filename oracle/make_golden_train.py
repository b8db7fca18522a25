"""Golden vectors of the teacher-forced forward pass (row f4): the UNMODIFIED reference's forward_train run here.

TEST INFRASTRUCTURE (build container only; needs /root/reference).  SurfaceFormer_Parallel.forward_train (model_para.py:99-171) in eval
mode (dropout off, scheduled_sampling_ratio = 0) on polygon batches with real labels, for the trained tiny and E = 512 checkpoints; then
Trainer.compute_loss's arithmetic (trainer.py:61-79: bmm, cross_entropy with ignore_index = PAD, token accuracy) with torch itself.
Stored per case: pointer (fp32 [N*F, T-1, E]; every `seq_step`-th sequence / `pos_step`-th position for the large models), the float64 evaluation of the same rows,
label, loss, accuracy, argmax predictions.

    python oracle/make_golden_train.py      # writes tests/golden/train_forward_*.npz
"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference")
sys.path.insert(0, os.path.join(ROOT, "oracle"))

from faceformer_b200 import synth  # noqa: E402
from faceformer_b200.config import MID, MODE_PARALLEL, MODE_SEQ2SEQ, SEQ2SEQ, TINY  # noqa: E402
from make_golden import build_reference  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")
CASES = {
    "train_forward_tiny": dict(cfg=TINY, n=5, weights="tiny_trained_parallel.npz", seed=41, seq_step=1),
    "train_forward_mid": dict(cfg=MID, n=6, weights="mid_trained_parallel.npz", seed=42, seq_step=5),
    # SurfaceFormer.forward_train (model.py:98-157): synthetic weights, crafted teacher sequences (synth.seq2seq_labels)
    "train_forward_seq2seq_tiny": dict(cfg=TINY, mode=MODE_SEQ2SEQ, n=4, weights=("synth", 9), seed=43, seq_step=1),
    "train_forward_seq2seq": dict(cfg=SEQ2SEQ, mode=MODE_SEQ2SEQ, n=2, weights=("synth", 10), seed=44, seq_step=1, pos_step=3),
}


def run(m, batch):
    tb = {k: torch.from_numpy(v) for k, v in batch.items()}
    with torch.no_grad():
        out = m.forward_train(tb)                                   # eval mode: dropout off; no scheduled sampling
        emb, ptr, labels = out["embedding"], out["pointer"], out["label"]
        logits = torch.bmm(emb, ptr.transpose(1, 2))                # trainer.py:64-66
        loss = torch.nn.functional.cross_entropy(logits, labels, ignore_index=0, reduction="sum")
        valid = labels != 0
        pred = torch.argmax(logits, dim=1)
        acc = float((valid * (pred == labels)).sum()) / (float(valid.sum()) + 1e-10)
        loss = float(loss / valid.sum())
    return ptr.numpy(), labels.numpy(), loss, acc, pred.numpy(), emb.numpy()


def main():
    torch.set_num_threads(os.cpu_count())
    for name, spec in CASES.items():
        cfg, mode = spec["cfg"], spec.get("mode", MODE_PARALLEL)
        if mode == MODE_PARALLEL:
            sd = synth.load_state_dict_npz(os.path.join(GOLDEN, spec["weights"]))
            batch = synth.polygon_batch(cfg, spec["n"], seed=spec["seed"])
        else:
            sd = synth.synth_state_dict(cfg, mode, spec["weights"][1], "diverse")
            batch = synth.seq2seq_labels(cfg, synth.synth_batch(cfg, mode, spec["n"], spec["seed"], lo=5, hi=min(60, cfg.num_lines)), spec["seed"])
        m = build_reference(cfg, mode, sd)
        ptr, labels, loss, acc, pred, emb = run(m, batch)
        memory = emb[::emb.shape[0] // spec["n"]]                     # [N, L, E]: one copy per wireframe, rows of padded edges included
        m64 = build_reference(cfg, mode, sd).double()
        b64 = {k: (v.astype(np.float64) if v.dtype == np.float32 else v) for k, v in batch.items()}
        ptr64, _, loss64, _, _, _ = run(m64, b64)
        st, ps = spec["seq_step"], spec.get("pos_step", 1)
        meta = dict(name=name, cfg=cfg.to_dict(), mode=mode, n=spec["n"], weights=spec["weights"], seed=spec["seed"], seq_step=st, pos_step=ps,
                    torch=torch.__version__)
        np.savez_compressed(os.path.join(GOLDEN, name + ".npz"), meta=json.dumps(meta), pointer=ptr[::st, ::ps].astype(np.float32),
                            pointer64=ptr64[::st, ::ps], label=labels, memory=memory.astype(np.float32), loss=np.float64(loss), loss64=np.float64(loss64), acc=np.float64(acc), pred=pred)
        print(f"{name}: pointer {ptr.shape} (kept {ptr[::st].shape}), loss {loss:.6f} (f64 {loss64:.6f}), token accuracy {acc:.4f}, max |pointer| {np.abs(ptr).max():.2f}")


if __name__ == "__main__":
    main()
