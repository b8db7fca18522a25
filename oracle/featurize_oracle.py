"""CPU restatement of the reference's input featurisation -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this module; the product path
(faceformer_b200.Engine.featurize -> ffb_featurize) never does.

Follows /root/reference/faceformer/datasets/data_para.py:
  sample_points            :8-11    2-point edge -> line, anything else -> curve
  sample_points_on_line    :14-19   x1 + (x2 - x1) * linspace(0, 1, P), float64
  sample_points_on_curve   :22-25   curve[linspace(0, n - 1, P).round(0).astype(int)]
  __getitem__              :59-68   input [num_lines, P, D] float32 zero-initialised, row i <- sample_points(edge_i);
                                    input_mask ones, first len(edges) cleared; num_input = len(edges)
Pinned by tests/golden/featurize.npz (outputs of the reference's own functions, oracle/make_golden_featurize.py).
"""
import numpy as np


def sample_points_on_line(line, num_samples):                       # data_para.py:14-19
    t = np.linspace(0, 1, num_samples)
    x1, y1, x2, y2 = line[0][0], line[0][1], line[1][0], line[1][1]
    x = x1 + (x2 - x1) * t
    y = y1 + (y2 - y1) * t
    return np.vstack([x, y]).T


def sample_points_on_curve(curve, num_samples):                     # data_para.py:22-25
    samples = np.linspace(0, len(curve) - 1, num_samples).round(0).astype(int)
    return np.array(curve)[samples]


def sample_points(edge, num_samples=50):                            # data_para.py:8-11
    if len(edge) == 2:
        return sample_points_on_line(edge, num_samples)
    return sample_points_on_curve(edge, num_samples)


def featurize(wireframes, num_lines, num_points_per_line=50, point_dim=2):
    """wireframes: list of edge lists.  -> (input [N, num_lines, P, D] f32, input_mask [N, num_lines] bool, num_input [N] i64)
    exactly as the default collate stacks the per-sample arrays of __getitem__ (data_para.py:59-68,98-107)."""
    n = len(wireframes)
    inp = np.zeros((n, num_lines, num_points_per_line, point_dim), dtype=np.float32)
    mask = np.ones((n, num_lines), dtype=bool)
    ni = np.zeros((n,), dtype=np.int64)
    for w, edges in enumerate(wireframes):
        for i, edge in enumerate(edges):
            inp[w, i, :num_points_per_line] = sample_points(edge, num_points_per_line)
        mask[w, :len(edges)] = 0
        ni[w] = len(edges)
    return inp, mask, ni


def synth_wireframes(n, num_lines, seed, lo=1):
    """Seeded polylines in the JSON's value range: 60 % segments, 40 % polylines of 1..300 points (incl. exact .5 index ties)."""
    rng = np.random.default_rng([seed, 7919])
    out = []
    for _ in range(n):
        ne = int(rng.integers(lo, num_lines + 1))
        edges = []
        for _ in range(ne):
            if rng.random() < 0.6:
                k = 2
            else:
                k = int(rng.choice([1, 3, 4, 50, 99, 148, 197, 246, int(rng.integers(3, 301))]))
            pts = rng.uniform(-1.0, 1.0, size=(k, 2))
            if rng.random() < 0.2:
                pts = np.round(pts, 3)                     # JSON files carry rounded decimals
            edges.append(pts.tolist())
        out.append(edges)
    return out
