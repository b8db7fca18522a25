"""Seeded cases for the set-level metrics of Trainer.face_accuracy (trainer.py:210-300) -- TEST INFRASTRUCTURE ONLY.

The expected values in tests/golden/metrics.npz come from the reference's OWN face_accuracy / parse_parallel_faces (executed from
the source text of faceformer/trainer.py, oracle/make_golden_metrics.py); this module only rebuilds the inputs from the seed.
"""
import numpy as np

from .faces_oracle import TOKEN_LEN, synth_case


def synth_metrics_case(n, num_lines, T, seed):
    """-> (raw_datas: list of dict(edges, pairings), predict int64 [n,F,T], label int64 [n,F,T]).  Labels are drawn like predictions
    (true loops in every rotation, multi-loop faces, broken loops ...) from a second stream, so that predictions and labels overlap
    partly; pairings map ~30 % of the co-edge indices onto a lower index (string keys, like the dataset JSON)."""
    wfs, pred = synth_case(n, num_lines, T, seed)
    wfs2, lab = synth_case(n, num_lines, T, seed)            # same wireframes (same stream) ...
    assert all(a == b for a, b in zip(wfs, wfs2))
    rng = np.random.default_rng([seed, 32452867])
    F = pred.shape[1]
    for w in range(n):                                       # ... labels: a shuffled half of the prediction rows + fresh rows of another seed
        perm = rng.permutation(F)
        lab[w] = pred[w][perm]
    _, other = synth_case(n, num_lines, T, seed + 1000)
    ne_other = other.shape[1]
    for w in range(n):
        ne = len(wfs[w])
        k = F // 2
        rows = other[w, :min(k, ne_other)].copy()
        rows[(rows >= TOKEN_LEN + ne)] = TOKEN_LEN            # keep label indices in range (the dataset never has out-of-range labels)
        lab[w, :rows.shape[0]] = rows
    raw = []
    for w in range(n):
        ne = len(wfs[w])
        pairings = {}
        for i in range(1, ne):
            if rng.random() < 0.3:
                pairings[str(i)] = int(rng.integers(0, i))
        raw.append({"edges": wfs[w], "pairings": pairings, "dominant_directions": [[1.0, 0.0, 0.0], [0.0, 1.0, 0.0], [0.0, 0.0, 1.0]]})
    return raw, pred, lab
