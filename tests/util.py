"""Shared helpers for the tests: rebuild a golden case's inputs/weights from its seeds."""
import json
import os

import numpy as np

from faceformer_b200 import synth
from faceformer_b200.config import ModelConfig

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
LOGIT_TOL = 1e-4          # north_star: "fp32 pointer logits within 1e-4"
LOGIT_SCALE = 32.0        # the logit magnitude that budget was derived at (std 5.7, max 37)


def load_case(name):
    """-> dict(cfg, mode, sd, batch, predict, steps, last_logits, memory, prefix, prefix_logits)."""
    with np.load(os.path.join(GOLDEN, name + ".npz")) as z:
        g = {k: z[k] for k in z.files}
    meta = json.loads(str(g.pop("meta")))
    cfg = ModelConfig(**meta["cfg"])
    mode, n = meta["mode"], meta["n"]
    w = meta["weights"]
    if w[0] == "synth":
        sd = synth.synth_state_dict(cfg, mode, seed=w[1], recipe=w[2])
    else:
        sd = synth.load_state_dict_npz(os.path.join(GOLDEN, w[1]))
    i = meta["inputs"]
    if i[0] == "synth":
        ne = None if i[2] is None else np.asarray(i[2], np.int64)
        batch = synth.synth_batch(cfg, mode, n, seed=i[1], num_edges=ne)
    else:
        batch = synth.polygon_batch(cfg, n, seed=i[1])
    g.update(cfg=cfg, mode=mode, sd=sd, batch=batch, meta=meta, steps=int(g["steps"]))
    return g


def valid_rows_mask(batch, cfg):
    """bool [N, L]: True for un-masked memory rows (4 token rows + valid edges)."""
    m = ~np.asarray(batch["input_mask"], bool)
    return np.concatenate([np.ones((m.shape[0], cfg.num_token), bool), m], axis=1)


def logits_close(a, b, tol=LOGIT_TOL, b64=None):
    """Masked entries must be exactly finfo.min on both sides; the rest within tol of the reference's fp32 logits `b`.

    tol is ABSOLUTE (1e-4) while max|logit| <= 32.  The trained fixture produces logits up to ~210,
    where one fp32 ulp is already 1.5e-5 and the reference's own fp32-vs-fp64 noise exceeds 1e-4,
    so beyond 32 the budget grows proportionally: tol * max|logit| / 32 (3.1e-6 relative).

    Noise-floor clause (b64 = the reference evaluated in float64, stored in the fixture): two correct fp32 evaluations of
    the same function can be farther apart than 1e-4 when the reference's own fp32 result is ~1e-4 from the exact one
    (measured 0.8e-4 on the ours.yml-size fixtures, profiles/logit_noise.py).  Such a comparison passes if our logits are no
    farther from the float64 evaluation than NOISE_FACTOR x the reference's own fp32 logits are (NOISE_FACTOR = 1: at least as close to
    the exact result as the reference itself)."""
    a, b = np.asarray(a), np.asarray(b)
    fmin = np.finfo(np.float32).min
    ma, mb = a == fmin, b == fmin
    if not np.array_equal(ma, mb):
        return False, float("inf")
    if not (~ma).any():
        return True, 0.0
    d = float(np.max(np.abs(a[~ma] - b[~mb])))
    scale = max(1.0, float(np.max(np.abs(b[~mb]))) / LOGIT_SCALE)
    if d <= tol * scale:
        return True, d
    if b64 is not None:
        t = np.asarray(b64, np.float64)[~mb]
        n32 = float(np.max(np.abs(b[~mb].astype(np.float64) - t)))
        d64 = float(np.max(np.abs(a[~ma].astype(np.float64) - t)))
        return d64 <= NOISE_FACTOR * n32, d
    return False, d


NOISE_FACTOR = 1.0        # r2 (float64 encoder + folded float64 head): 0.3x - 0.9x over all fixtures and pipelines

CASES_ALL = ["tiny_parallel_trained", "tiny_parallel_trained_b", "tiny_parallel_ragged", "tiny_seq2seq",
             "ours_parallel_small", "seq2seq_single64",
             "perspective_small",     # BASELINE.json configs[3] geometry (ours-perspective.yml), greedy
             "ours_wide300",          # configs[1] "<= 512 edges": 274 memory rows (> 256: mma.sync cross-attention)
             "mid_parallel_trained",  # E = 512 / H = 8 TRAINED checkpoint (non-degenerate weights on the tcgen05 grid), 16 and 24 wireframes
             "mid_parallel_trained_b"]
