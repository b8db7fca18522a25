"""Beam-search decode on top of the oracle's loop body -- the SPECIFICATION of BASELINE.json configs[3] ("beam=4").

TEST INFRASTRUCTURE.  PARITY UNPINNED: the reference has no beam search (grep -ri beam /root/reference -> 0 hits), so there is
nothing to pin this against except its own greedy loop: beam = 1 must reproduce SurfaceFormer_Parallel.forward_eval
(model_para.py:181-241) token for token, which tests/test_beam.py checks against the reference-generated goldens.

Specification (parallel mode only; one search per anchor sequence, W = beam width):
  * every anchor owns W hypotheses (token prefixes of equal length) with cumulative scores `cum` in float64;
    initially hypothesis 0 = [anchor] with cum 0, hypotheses 1..W-1 are dead (cum = -inf);
  * step: pointer logits [L] of every hypothesis exactly as the greedy loop computes them (model_para.py:217-227;
    masked rows = finfo(float32).min); logp = logit - logsumexp(logit over UN-MASKED rows), evaluated in float64;
  * per hypothesis the W best rows by (logit descending, row index ascending) are its candidates -- for W = 1 this is
    torch.argmax's first maximum (model_para.py:179);
  * the candidates of all LIVE hypotheses are merged and the W best by (cum + logp descending, hypothesis index
    ascending, candidate rank ascending) become the new hypotheses 0..W-1 (so hypothesis 0 is always the best one);
  * the loop runs T-1 steps; like the reference it has no per-sequence termination, and it stops early when the newly
    emitted tokens of ALL hypotheses of ALL anchors are special tokens (< 4), the beam analogue of model_para.py:232;
  * output: predict [N,F,T] = hypothesis 0 of every anchor (zero padded after an early stop, model_para.py:236),
    plus all beams [N,F,W,T] and their scores [N,F,W].
"""
from __future__ import annotations

import numpy as np

from . import faceformer_oracle as orc

F32_MIN = np.finfo(np.float32).min


def beam_select(logits, cum, W):
    """One selection step for ONE anchor.  logits f32 [W, L] (masked rows == finfo.min), cum f64 [W] (-inf = dead).
    Returns (parent [W] int, token [W] int, new_cum [W] f64)."""
    cands = []
    for w in range(W):
        if not np.isfinite(cum[w]):
            continue
        lg = logits[w]
        valid = lg != F32_MIN
        x = lg[valid].astype(np.float64)
        m = x.max()
        lse = m + np.log(np.exp(x - m).sum())
        order = np.lexsort((np.arange(lg.shape[0]), -lg.astype(np.float64)))       # logit desc, index asc
        order = [int(i) for i in order if valid[i]][:W]
        for r, tok in enumerate(order):
            cands.append((-(cum[w] + (float(lg[tok]) - lse)), w, r, tok))
    cands.sort()
    assert len(cands) >= W, "fewer candidates than beams"
    parent = np.array([c[1] for c in cands[:W]], np.int64)
    token = np.array([c[3] for c in cands[:W]], np.int64)
    new_cum = np.array([-c[0] for c in cands[:W]], np.float64)
    return parent, token, new_cum


def forward_eval_beam(sd, cfg, inputs, W, return_trace=False, max_steps=None):
    """Beam-W decode of SurfaceFormer_Parallel (see the module docstring).  W = 1 is the greedy loop."""
    memory, input_mask, pos, qpos = orc.encode(sd, cfg, orc.MODE_PARALLEL, inputs)
    N = memory.shape[1]
    num_token, T = cfg["num_token"], cfg["max_face_length"]
    num_input = np.asarray(inputs["num_input"]).astype(np.int64)
    Fm = int(num_input.max())
    anchors = np.tile(np.arange(Fm, dtype=np.int64), (N, 1))
    for i, ne in enumerate(num_input):
        anchors[i, int(ne):] = num_token - 1
    B = N * Fm
    hyp = np.repeat(anchors.reshape(1, B, 1), W, axis=2)                    # [P, B, W]
    cum = np.full((B, W), -np.inf)
    cum[:, 0] = 0.0
    mem_rep = np.repeat(memory, Fm * W, axis=1)                             # sequence index = b * W + w
    mask_rep = np.repeat(input_mask, Fm * W, axis=0)
    steps, trace = 0, []
    for step in range(T - 1 if max_steps is None else min(T - 1, max_steps)):
        P = hyp.shape[0]
        _, logits, _ = orc.decode_step(sd, cfg, mem_rep, mask_rep, pos, qpos, hyp.reshape(P, B * W))
        logits = logits.reshape(B, W, -1)
        new_hyp = np.zeros((P + 1, B, W), np.int64)
        for b in range(B):
            parent, token, new_cum = beam_select(logits[b], cum[b], W)
            new_hyp[:P, b, :] = hyp[:, b, parent]
            new_hyp[P, b, :] = token
            cum[b] = new_cum
        hyp = new_hyp
        steps += 1
        if return_trace:
            trace.append(logits)
        if np.all(hyp[-1] < num_token):
            break
    pad = np.zeros((T - hyp.shape[0], B, W), np.int64)
    hyp = np.concatenate([hyp, pad], 0)
    beams = hyp.transpose(1, 2, 0).reshape(N, Fm, W, T)
    out = {"predict": np.ascontiguousarray(beams[:, :, 0]), "beams": beams, "scores": cum.reshape(N, Fm, W), "steps": steps}
    if return_trace:
        out["logits"] = trace
    return out
