"""Generate golden vectors by running the UNMODIFIED reference here.

TEST INFRASTRUCTURE (build container only; needs /root/reference).  The
reference ships no golden vectors for this path, so the
fixtures under ``tests/golden/`` are outputs of the reference model itself:
``SurfaceFormer_Parallel`` / ``SurfaceFormer`` imported from /root/reference,
strict-loaded with seeded synthetic weights (``faceformer_b200.synth``) or the
trained tiny checkpoint (``oracle/train_fixture.py``), evaluated on seeded
synthetic batches with ``model(batch)`` -- the same call ``Trainer.forward``
makes (trainer.py:27-28).

Each fixture stores only seeds + expected outputs (inputs and weights are
regenerated from the seed on the GPU box, where /root/reference does not exist):
  predict          int64  the reference's inputs['predict']
  steps            int    executed decode steps S
  last_logits      f32    masked pointer logits [B,L] of the last executed step
  memory           f32    encoder memory [N,L,E] (padded rows included, as the reference computes them)
  prefix / prefix_logits   a random forced token prefix [P,B] and its logits [B,L]
                   through the reference's own sub-modules (model_para.py:217-227)
  last_logits64 / prefix_logits64   the same logits with the reference model run in float64 (noise floor of fp32)

    python oracle/make_golden.py            # writes tests/golden/*.npz
"""
from __future__ import annotations

import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference")

from faceformer_b200 import synth  # noqa: E402
from faceformer_b200.config import (MID, MODE_PARALLEL, MODE_SEQ2SEQ, OURS, OURS_PERSPECTIVE, SEQ2SEQ, TINY, ModelConfig)  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")

# name -> spec.  weights: ("synth", seed, recipe) or ("file", npz name)
# inputs:  ("synth", seed, num_edges list | None, lo, hi) or ("polygon", seed)
CASES = {
    "tiny_parallel_trained": dict(cfg=TINY, mode=MODE_PARALLEL, n=4, weights=("file", "tiny_trained_parallel.npz"),
                                  inputs=("polygon", 11), prefix_P=5),
    "tiny_parallel_trained_b": dict(cfg=TINY, mode=MODE_PARALLEL, n=7, weights=("file", "tiny_trained_parallel.npz"),
                                    inputs=("polygon", 12), prefix_P=9),
    # E = 512 / H = 8 checkpoint TRAINED with the reference's forward_train (oracle/train_fixture.py --cfg mid): non-degenerate weights on the
    # tcgen05 grid, 16 / 24 polygon wireframes (diverse tokens, early stop)
    "mid_parallel_trained": dict(cfg=MID, mode=MODE_PARALLEL, n=16, weights=("file", "mid_trained_parallel.npz"),
                                 inputs=("polygon", 21), prefix_P=6),
    "mid_parallel_trained_b": dict(cfg=MID, mode=MODE_PARALLEL, n=24, weights=("file", "mid_trained_parallel.npz"),
                                   inputs=("polygon", 22), prefix_P=8),
    "tiny_parallel_ragged": dict(cfg=TINY, mode=MODE_PARALLEL, n=5, weights=("synth", 3, "diverse"),
                                 inputs=("synth", 5, [1, 28, 3, 17, 9]), prefix_P=7),
    "tiny_seq2seq": dict(cfg=TINY, mode=MODE_SEQ2SEQ, n=3, weights=("synth", 4, "diverse"),
                         inputs=("synth", 6, [5, 28, 12]), prefix_P=11),
    "ours_parallel_small": dict(cfg=OURS, mode=MODE_PARALLEL, n=2, weights=("synth", 0, "diverse"),
                                inputs=("synth", 0, [10, 14]), prefix_P=6),
    "seq2seq_single64": dict(cfg=SEQ2SEQ, mode=MODE_SEQ2SEQ, n=1, weights=("synth", 1, "diverse"),
                             inputs=("synth", 1, [64]), prefix_P=20),
    # BASELINE.json configs[3] geometry (ours-perspective.yml: num_lines 202, T 38); greedy only -- the reference has no beam search
    "perspective_small": dict(cfg=OURS_PERSPECTIVE, mode=MODE_PARALLEL, n=3, weights=("synth", 2, "diverse"),
                              inputs=("synth", 2, [12, 30, 21]), prefix_P=8),
    # BASELINE.json configs[1] "<= 512 edges": a wireframe with more than 256 memory rows (274), num_lines overridden to 300
    "ours_wide300": dict(cfg=OURS.replace(num_lines=300), mode=MODE_PARALLEL, n=1, weights=("synth", 5, "diverse"),
                         inputs=("synth", 7, [270]), prefix_P=6),
}

# encoder-only fixture for BASELINE.json configs[4] (2048-edge wireframes): every 8th memory row of the reference's encoder output
ENCODER_CASES = {
    "encoder_2048": dict(cfg=OURS.replace(num_lines=2048), mode=MODE_PARALLEL, n=1, weights=("synth", 6, "diverse"),
                         inputs=("synth", 8, [2048]), row_step=8),
}


def load_weights(spec, cfg, mode):
    if spec[0] == "synth":
        return synth.synth_state_dict(cfg, mode, seed=spec[1], recipe=spec[2])
    return synth.load_state_dict_npz(os.path.join(GOLDEN, spec[1]))


def load_inputs(spec, cfg, mode, n):
    if spec[0] == "synth":
        ne = None if spec[2] is None else np.asarray(spec[2], np.int64)
        return synth.synth_batch(cfg, mode, n, seed=spec[1], num_edges=ne)
    return synth.polygon_batch(cfg, n, seed=spec[1])


def build_reference(cfg: ModelConfig, mode, sd):
    from faceformer.models import SurfaceFormer, SurfaceFormer_Parallel
    cls = SurfaceFormer_Parallel if mode == MODE_PARALLEL else SurfaceFormer
    m = cls(**cfg.model_kwargs(mode)).eval()
    m.load_state_dict({k: torch.from_numpy(np.asarray(v)) for k, v in sd.items()}, strict=True)
    return m


def reference_logits(m, mode, batch, prefix):
    """Loop body of forward_eval on a given prefix, through the reference's sub-modules
    (model_para.py:191-227 / model.py:177-203).  Returns (memory [N,L,E], logits [B,L])."""
    tb = {k: torch.from_numpy(v) for k, v in batch.items()}
    with torch.no_grad():
        input_mask = m.process_masks(tb["input_mask"])
        val, pos, qpos = m.get_embeddings(tb["input"], tb["label"])
        source, pos = m.patch_source(val, pos)
        qpos = qpos.transpose(0, 1)
        memory = m.encoder(source, src_key_padding_mask=input_mask, pos=pos)
        mem_nle = memory.transpose(0, 1).contiguous().numpy()
        if mode == MODE_PARALLEL:
            F = int(max(tb["num_input"]))
            memory = memory.repeat_interleave(F, 1)
            input_mask = input_mask.repeat_interleave(F, 0)
        pre = torch.from_numpy(prefix)
        tgt = torch.gather(memory, 0, pre.unsqueeze(-1).repeat(1, 1, m.num_model))
        ptr = m.project(m.decoder(tgt, memory, memory_key_padding_mask=input_mask, pos=pos, query_pos=qpos[:pre.size(0)]))
        emb = memory.transpose(0, 1)
        logit = torch.bmm(emb, ptr.permute(1, 2, 0)[..., -1:])
        logit = logit.masked_fill(input_mask.unsqueeze(-1), torch.finfo(logit.dtype).min)
    return mem_nle, logit[..., 0].numpy()


def random_prefix(cfg, mode, batch, P, seed):
    """Random valid tokens: row indices of un-masked memory rows of the owning wireframe."""
    rng = np.random.default_rng([seed, 32452843])
    nvalid = (~batch["input_mask"]).sum(1) + cfg.num_token
    if mode == MODE_PARALLEL:
        F = int(batch["num_input"].max())
        owner = np.repeat(np.arange(len(nvalid)), F)
    else:
        owner = np.arange(len(nvalid))
    return np.stack([rng.integers(0, nvalid[owner]) for _ in range(P)]).astype(np.int64)


def make_case(name, spec):
    cfg, mode, n = spec["cfg"], spec["mode"], spec["n"]
    sd = load_weights(spec["weights"], cfg, mode)
    batch = load_inputs(spec["inputs"], cfg, mode, n)
    m = build_reference(cfg, mode, sd)
    torch.set_num_threads(os.cpu_count())
    t0 = time.time()
    with torch.no_grad():
        out = m({k: torch.from_numpy(v) for k, v in batch.items()})
    dt = time.time() - t0
    predict = out["predict"].numpy()
    # executed steps: length of the non-padded part is not recoverable from zeros alone -> re-derive with the stop rule
    T = cfg.seq_len(mode)
    flat = predict.reshape(-1, T)
    if mode == MODE_PARALLEL:
        steps = T - 1
        for s in range(1, T):
            if np.all(flat[:, s] < cfg.num_token):
                steps = s
                break
    else:
        steps, eos = T - 1, 0
        for s in range(1, T):
            eos += int((flat[:, s] == 3).sum())
            if eos == flat.shape[0]:
                steps = s
                break
    # last executed step's logits = forced prefix of the first `steps` rows
    _, last_logits = reference_logits(m, mode, batch, np.ascontiguousarray(flat[:, :steps].T))
    assert np.array_equal(last_logits.argmax(1), flat[:, steps]), "stop-rule reconstruction is inconsistent"
    prefix = random_prefix(cfg, mode, batch, spec["prefix_P"], seed=len(name))
    memory, prefix_logits = reference_logits(m, mode, batch, prefix)
    # the same two evaluations with the reference model in float64: the distance between the reference's own fp32 result
    # and the exact one is the noise floor any fp32 implementation is judged against (tests/util.py logits_close)
    m64 = build_reference(cfg, mode, sd).double()
    b64 = {k: (v.astype(np.float64) if v.dtype == np.float32 else v) for k, v in batch.items()}
    _, last_logits64 = reference_logits(m64, mode, b64, np.ascontiguousarray(flat[:, :steps].T))
    _, prefix_logits64 = reference_logits(m64, mode, b64, prefix)
    meta = dict(name=name, cfg=cfg.to_dict(), mode=mode, n=n, weights=list(spec["weights"]),
                inputs=[x if not isinstance(x, np.ndarray) else x.tolist() for x in spec["inputs"]],
                torch=torch.__version__, threads=torch.get_num_threads(), ref_seconds=round(dt, 3))
    np.savez_compressed(os.path.join(GOLDEN, name + ".npz"), meta=json.dumps(meta), predict=predict,
                        steps=np.int64(steps), last_logits=last_logits, memory=memory.astype(np.float32),
                        prefix=prefix, prefix_logits=prefix_logits, last_logits64=last_logits64, prefix_logits64=prefix_logits64)
    print(f"{name}: predict {predict.shape} steps {steps} distinct {len(np.unique(predict))} ref {dt:.2f}s", flush=True)


def make_encoder_case(name, spec):
    """memory = encoder(embeddings) through the reference's own sub-modules (model_para.py:191-210), rows subsampled."""
    cfg, mode, n = spec["cfg"], spec["mode"], spec["n"]
    sd = load_weights(spec["weights"], cfg, mode)
    batch = load_inputs(spec["inputs"], cfg, mode, n)
    m = build_reference(cfg, mode, sd)
    torch.set_num_threads(os.cpu_count())
    tb = {k: torch.from_numpy(v) for k, v in batch.items()}
    t0 = time.time()
    with torch.no_grad():
        input_mask = m.process_masks(tb["input_mask"])
        val, pos, _ = m.get_embeddings(tb["input"], tb["label"])
        source, pos = m.patch_source(val, pos)
        memory = m.encoder(source, src_key_padding_mask=input_mask, pos=pos).transpose(0, 1).contiguous().numpy()
    dt = time.time() - t0
    rows = np.arange(0, memory.shape[1], spec["row_step"])
    meta = dict(name=name, cfg=cfg.to_dict(), mode=mode, n=n, weights=list(spec["weights"]),
                inputs=[x if not isinstance(x, np.ndarray) else x.tolist() for x in spec["inputs"]],
                torch=torch.__version__, threads=torch.get_num_threads(), ref_seconds=round(dt, 3))
    np.savez_compressed(os.path.join(GOLDEN, name + ".npz"), meta=json.dumps(meta), rows=rows, memory_rows=memory[:, rows].astype(np.float32),
                        steps=np.int64(0))
    print(f"{name}: memory {memory.shape} -> {len(rows)} rows kept, ref {dt:.2f}s", flush=True)


def main():
    only = sys.argv[1:]
    for name, spec in CASES.items():
        if only and name not in only:
            continue
        make_case(name, spec)
    for name, spec in ENCODER_CASES.items():
        if only and name not in only:
            continue
        make_encoder_case(name, spec)


if __name__ == "__main__":
    main()
