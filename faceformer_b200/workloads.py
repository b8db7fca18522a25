"""bench.py workloads beyond the headline batch: the other BASELINE.json `configs`, each measured to the same contract
(device-timed `value`, host-buffer `e2e`, `roofline`).  Host-side plumbing only: every timed call goes through the C ABI.

  ours_n1        configs/ours.yml, ONE wireframe per model(batch) call -- the reference's own test loop (trainer.py:51 forces batch 1)
  seq2seq_n1_64  BASELINE configs[0]: configs/seq2seq.yml, one 64-edge wireframe, 258 greedy steps
  split507       BASELINE configs[2]: 507 synthetic wireframes (5 % of the 10 124 ids, split_jsons.py:13-17), global batch 128, every
                 global batch SPLIT over all ranks (strong scaling), predictions all-gathered
  beam4          BASELINE configs[3]: configs/ours-perspective.yml geometry, batch 64, beam width 4 (specified in oracle/beam_oracle.py)
  encoder2048    BASELINE configs[4]: encoder only, 2048-edge wireframes (L = 2052), tensor-core encoder mode
"""
from __future__ import annotations

import os

import numpy as np

from . import synth
from .config import MODE_PARALLEL, MODE_SEQ2SEQ, OURS, OURS_PERSPECTIVE, SEQ2SEQ


def decoder_step_bytes(cfg, vlen_sum, n_seq):
    """Algorithmic HBM bytes of ONE decode step (SURVEY.md 8d "Algorithmic bytes"): the decoder weights once (fp16x2 operand pairs = 4 B per
    parameter, shared by all sequences), the cross-attention K / V cache of the batch's wireframes (fp16x2, 4 B per element), the folded
    pointer-head rows (fp32) and the token ids.  Activations are on-chip in the ideal and not counted."""
    E, FF, Ld = cfg.num_model, cfg.num_feedforward, cfg.num_decoder_layers
    w = Ld * (4 * E * E + E * E + E * E + 2 * E * FF) + 0          # self in-proj 3E^2 + out E^2, cross q E^2 + out E^2, FFN 2*E*FF
    w_bytes = 4 * w
    kv_bytes = 4 * 2 * Ld * vlen_sum * E
    head_bytes = 4 * vlen_sum * (E + 4)
    return w_bytes + kv_bytes + head_bytes + 8 * n_seq


def encoder_flops(cfg, n_rows_list):
    """per wireframe 2*[L*6*(4E^2 + 2*E*FF) + 6*2*L^2*E + n*(in*E + E^2)] (104.7 GFLOP at L = 2052); the once-per-wireframe
    cross-attention K / V projections belong to the decode and are not part of the encoder-only workload."""
    E, FF, Le, ind = cfg.num_model, cfg.num_feedforward, cfg.num_encoder_layers, cfg.in_dim
    tot = 0.0
    for L in n_rows_list:
        n = L - cfg.num_token
        tot += 2.0 * (L * Le * (4 * E * E + 2 * E * FF) + Le * 2 * L * L * E + n * (ind * E + E * E))
    return tot


def make_small(name, args, torch, dev):
    """-> dict(cfg, mode, sd, calls=[(coords, mask, num_input) device tensors], host_calls=[numpy versions])"""
    if name == "seq2seq_n1_64":
        cfg, mode = SEQ2SEQ, MODE_SEQ2SEQ
        sd = synth.synth_state_dict(cfg, mode, 1, "diverse")                 # the seed of the seq2seq_single64 golden: no early EOS, 258 steps
        batches = [synth.synth_batch(cfg, mode, 1, seed=1, num_edges=np.array([64], np.int64))]
    else:
        cfg, mode = OURS, MODE_PARALLEL
        sd = synth.synth_state_dict(cfg, mode, args.seed, "diverse")
        ne = synth.synth_num_edges(cfg, 16, args.seed)                       # 16 wireframes, n ~ U[24, 216], one model(batch) call each
        batches = [synth.synth_batch(cfg, mode, 1, seed=args.seed + 100 + i, num_edges=np.array([n], np.int64)) for i, n in enumerate(ne)]
    calls, host_calls = [], []
    for b in batches:
        c = np.ascontiguousarray(b["input"].reshape(1, cfg.num_lines, -1))
        m = b["input_mask"].astype(np.uint8)
        ni = b["num_input"] if mode == MODE_PARALLEL else None
        host_calls.append((c, m, ni))
        calls.append((torch.from_numpy(c).to(dev), torch.from_numpy(m).to(dev), None if ni is None else torch.from_numpy(ni).to(dev)))
    return dict(cfg=cfg, mode=mode, sd=sd, calls=calls, host_calls=host_calls, batches=batches)
