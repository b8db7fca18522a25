"""Seeded synthetic wireframes and weights (numpy only — no torch, no reference).

The reference's checkpoints and datasets are Google-Drive downloads
(/root/reference/README.md:33,38) and are not available offline, so every
parity test, golden fixture and bench run uses the generators below.  They are
deterministic functions of (config, seed): a fixture only has to store the seed
and the expected outputs.

* ``synth_state_dict`` produces a dict with exactly the reference's
  ``state_dict`` names and shapes so it can be loaded
  strictly into ``SurfaceFormer`` / ``SurfaceFormer_Parallel``.
* ``synth_batch`` produces the dict a DataLoader would hand to
  ``Trainer.forward`` (data_para.py:56-110 after default collation).
"""
from __future__ import annotations

import numpy as np

from .config import MODE_PARALLEL, MODE_SEQ2SEQ, ModelConfig


# --------------------------------------------------------------------------- weights
def _xavier(rng, shape, gain=1.0):
    fan_out, fan_in = shape[0], shape[1]
    a = gain * np.sqrt(6.0 / (fan_in + fan_out))
    return rng.uniform(-a, a, size=shape).astype(np.float32)


def _bias(rng, n, fan_in):
    b = 1.0 / np.sqrt(fan_in)
    return rng.uniform(-b, b, size=(n,)).astype(np.float32)


def state_dict_names(cfg: ModelConfig, mode: int):
    """Ordered (name, shape, dtype) triples of the reference ``state_dict``."""
    E, FF, L, T = cfg.num_model, cfg.num_feedforward, cfg.mem_len, cfg.seq_len(mode)
    out = [("val_enc.embedding_token.weight", (cfg.num_token, E), "f4"),
           ("val_enc.embedding_value.0.weight", (E, cfg.in_dim), "f4"),
           ("val_enc.embedding_value.0.bias", (E,), "f4"),
           ("val_enc.embedding_value.2.weight", (E, E), "f4"),
           ("val_enc.embedding_value.2.bias", (E,), "f4"),
           ("pos_enc.position", (1, L), "i8"),
           ("pos_enc.pos_embed.weight", (L, E), "f4"),
           ("query_pos_enc.position", (1, T), "i8"),
           ("query_pos_enc.pos_embed.weight", (T, E), "f4")]

    def attn(prefix):
        return [(prefix + ".in_proj_weight", (3 * E, E), "f4"),
                (prefix + ".in_proj_bias", (3 * E,), "f4"),
                (prefix + ".out_proj.weight", (E, E), "f4"),
                (prefix + ".out_proj.bias", (E,), "f4")]

    def ffn_norms(prefix, n_norm):
        r = [(prefix + ".linear1.weight", (FF, E), "f4"), (prefix + ".linear1.bias", (FF,), "f4"),
             (prefix + ".linear2.weight", (E, FF), "f4"), (prefix + ".linear2.bias", (E,), "f4")]
        for i in range(1, n_norm + 1):
            r += [(f"{prefix}.norm{i}.weight", (E,), "f4"), (f"{prefix}.norm{i}.bias", (E,), "f4")]
        return r

    for l in range(cfg.num_encoder_layers):
        p = f"encoder.layers.{l}"
        out += attn(p + ".self_attn") + ffn_norms(p, 2)
    out += [("encoder.norm.weight", (E,), "f4"), ("encoder.norm.bias", (E,), "f4")]
    for l in range(cfg.num_decoder_layers):
        p = f"decoder.layers.{l}"
        out += attn(p + ".self_attn") + attn(p + ".multihead_attn") + ffn_norms(p, 3)
    out += [("decoder.norm.weight", (E,), "f4"), ("decoder.norm.bias", (E,), "f4"),
            ("project.weight", (E, E), "f4"), ("project.bias", (E,), "f4")]
    return out


def synth_state_dict(cfg: ModelConfig, mode: int, seed: int = 0, recipe: str = "diverse"):
    """Random weights with the reference's names/shapes.

    recipe "init":    the reference constructor's distributions — xavier-uniform on
                      every >1-d parameter (model_para.py:50-53), torch Linear-style
                      uniform biases, zero attention biases, unit LayerNorms.
    recipe "diverse": same matrices, but LayerNorm gains/biases, attention biases and
                      the pointer projection are perturbed so that greedy decode does
                      not collapse to a single token (SURVEY.md section 7
                      "Degenerate synthetic weights").  This is the default for
                      parity tests and for the bench.
    """
    rng = np.random.default_rng(seed)
    sd = {}
    for name, shape, dt in state_dict_names(cfg, mode):
        if dt == "i8":
            sd[name] = np.arange(shape[1], dtype=np.int64)[None, :]
        elif len(shape) == 2:
            sd[name] = _xavier(rng, shape)
        elif name.endswith("in_proj_bias") or name.endswith("out_proj.bias"):
            sd[name] = np.zeros(shape, np.float32)
        elif ".norm" in name and name.endswith("weight"):
            sd[name] = np.ones(shape, np.float32)
        elif ".norm" in name and name.endswith("bias"):
            sd[name] = np.zeros(shape, np.float32)
        else:  # Linear biases
            fan_in = {"val_enc.embedding_value.0.bias": cfg.in_dim,
                      "val_enc.embedding_value.2.bias": cfg.num_model,
                      "project.bias": cfg.num_model}.get(name)
            if fan_in is None:
                fan_in = cfg.num_feedforward if "linear2" in name else cfg.num_model
            sd[name] = _bias(rng, shape[0], fan_in)
    if recipe == "init":
        return sd
    if recipe != "diverse":
        raise ValueError(f"unknown recipe {recipe!r}")
    for name in list(sd):
        v = sd[name]
        if ".norm" in name and name.endswith("weight"):
            sd[name] = (v + rng.normal(0, 0.25, v.shape)).astype(np.float32)
        elif ".norm" in name and name.endswith("bias"):
            sd[name] = rng.normal(0, 0.1, v.shape).astype(np.float32)
        elif name.endswith("in_proj_bias") or name.endswith("out_proj.bias"):
            sd[name] = rng.normal(0, 0.05, v.shape).astype(np.float32)
        elif name.endswith("in_proj_weight"):
            # sharper attention: scale the q/k blocks
            E = cfg.num_model
            v = v.copy()
            v[: 2 * E] *= 2.0
            sd[name] = v
        elif name.endswith("pos_embed.weight") or name == "val_enc.embedding_token.weight":
            sd[name] = (v * 4.0).astype(np.float32)
        elif name == "val_enc.embedding_value.0.weight":
            sd[name] = (v * 3.0).astype(np.float32)
    return sd


# --------------------------------------------------------------------------- inputs
def _sample_points(rng, cfg: ModelConfig):
    """One edge polyline resampled to P points, as data_para.py:8-25 does."""
    P = cfg.num_points_per_line
    if rng.random() < 0.7:   # straight segment: linspace between two endpoints (data_para.py:14-20)
        a, b = rng.uniform(-1, 1, 2), rng.uniform(-1, 1, 2)
        t = np.linspace(0, 1, P)
        return np.stack([a[0] + (b[0] - a[0]) * t, a[1] + (b[1] - a[1]) * t], axis=1)
    # polyline on a circular arc, index-resampled (data_para.py:22-25)
    npts = int(rng.integers(20, 201))
    c = rng.uniform(-0.5, 0.5, 2)
    r = rng.uniform(0.05, 0.5)
    t0, dt = rng.uniform(0, 2 * np.pi), rng.uniform(0.3, 2 * np.pi)
    th = t0 + dt * np.linspace(0, 1, npts)
    curve = np.stack([c[0] + r * np.cos(th), c[1] + r * np.sin(th)], axis=1)
    idx = np.linspace(0, npts - 1, P).round(0).astype(int)
    return np.clip(curve[idx], -1, 1)


def synth_num_edges(cfg: ModelConfig, n: int, seed: int, lo: int = 24, hi: int | None = None,
                    dist: str = "uniform"):
    """Edge counts n_i per wireframe (SURVEY.md section 8d "Synthetic inputs")."""
    rng = np.random.default_rng([seed, 7919])
    hi = cfg.num_lines if hi is None else min(hi, cfg.num_lines)
    lo = max(1, min(lo, hi))
    if dist == "uniform":
        return rng.integers(lo, hi + 1, size=n).astype(np.int64)
    if dist == "real":   # real-like preset: clip(lognormal(ln 60, 0.5), 12, num_lines)
        v = np.exp(rng.normal(np.log(60.0), 0.5, size=n))
        return np.clip(np.round(v), min(12, hi), hi).astype(np.int64)
    raise ValueError(dist)


def synth_batch(cfg: ModelConfig, mode: int, n: int, seed: int = 0, num_edges=None,
                lo: int = 24, hi: int | None = None, dist: str = "uniform"):
    """A collated batch dict of numpy arrays (keys as data_para.py:98-107 / data.py)."""
    if num_edges is None:
        num_edges = synth_num_edges(cfg, n, seed, lo, hi, dist)
    num_edges = np.asarray(num_edges, dtype=np.int64)
    assert num_edges.shape == (n,) and num_edges.max(initial=0) <= cfg.num_lines
    rng = np.random.default_rng([seed, 104729])
    inp = np.zeros((n, cfg.num_lines, cfg.num_points_per_line, cfg.point_dim), np.float32)
    mask = np.ones((n, cfg.num_lines), np.bool_)
    for i, ne in enumerate(num_edges):
        for e in range(int(ne)):
            inp[i, e] = _sample_points(rng, cfg).astype(np.float32)
        mask[i, :ne] = False
    T = cfg.seq_len(mode)
    if mode == MODE_PARALLEL:
        label = np.zeros((n, cfg.num_lines, T), np.int64)      # only shape/dtype consumed
    else:
        label = np.zeros((n, T), np.int64)
    batch = {"id": np.arange(n, dtype=np.int64), "input": inp, "input_mask": mask,
             "label": label, "num_input": num_edges}
    return batch


# --------------------------------------------------------------------------- polygon wireframes
def polygon_wireframe(rng, cfg: ModelConfig, max_faces: int = 5):
    """A wireframe of 2..max_faces disjoint convex polygons (3-6 straight edges each) with the
    global edge order shuffled.  Returns (edges [n,P,2] float32, faces: list of loops of edge
    indices in traversal order).  Used to train / exercise the non-degenerate fixture."""
    P = cfg.num_points_per_line
    polys, total = [], 0
    for _ in range(int(rng.integers(2, max_faces + 1))):
        k = int(rng.integers(3, 7))
        if total + k > cfg.num_lines:
            break
        c = rng.uniform(-0.6, 0.6, 2)
        r = rng.uniform(0.1, 0.4)
        ang = np.sort(rng.uniform(0, 2 * np.pi, k))
        polys.append(c[None] + r * np.stack([np.cos(ang), np.sin(ang)], 1))
        total += k
    perm = rng.permutation(total)
    edges = np.zeros((total, P, 2), np.float32)
    faces = []
    t = np.linspace(0, 1, P)[:, None]
    g = 0
    for v in polys:
        loop = []
        for j in range(len(v)):
            a, b = v[j], v[(j + 1) % len(v)]
            idx = int(perm[g]); g += 1
            edges[idx] = (a[None] + (b - a)[None] * t).astype(np.float32)     # data_para.py:14-20
            loop.append(idx)
        faces.append(loop)
    return edges, faces


def polygon_sample(rng, cfg: ModelConfig, max_faces: int = 5):
    """One dataset item built like ABCDataset_Parallel.__getitem__ (data_para.py:56-110)."""
    edges, faces = polygon_wireframe(rng, cfg, max_faces)
    n, T = len(edges), cfg.max_face_length
    inp = np.zeros((cfg.num_lines, cfg.num_points_per_line, cfg.point_dim), np.float32)
    inp[:n] = edges
    mask = np.ones(cfg.num_lines, np.bool_); mask[:n] = False
    label = np.zeros((cfg.num_lines, T), np.int64)                   # token.PAD = 0
    ind = 0
    for loop in faces:                                               # one loop per face, type Plane(0)+offset
        for i in range(len(loop)):
            seq = np.roll(loop, i).tolist()                          # data_para.py:84
            label[ind, :len(seq)] = np.asarray(seq) + cfg.num_token  # data_para.py:91
            label[ind, len(seq)] = 1                                 # data_para.py:92
            ind += 1
    for i in range(ind, cfg.num_lines):
        label[i, 0] = cfg.num_token - 1                              # data_para.py:95
    return dict(input=inp, input_mask=mask, label=label, label_mask=(label == 0),
                num_input=np.int64(n))


def polygon_batch(cfg: ModelConfig, n: int, seed: int = 0, max_faces: int = 5):
    """Collated batch of polygon wireframes (parallel mode)."""
    rng = np.random.default_rng([seed, 15485863])
    items = [polygon_sample(rng, cfg, max_faces) for _ in range(n)]
    out = {k: np.stack([np.asarray(it[k]) for it in items]) for k in items[0]}
    out["id"] = np.arange(n, dtype=np.int64)
    return out


# --------------------------------------------------------------------------- checkpoint fixtures
def seq2seq_labels(cfg: ModelConfig, batch: dict, seed: int = 0):
    """Teacher sequences for a seq2seq batch (the layout of datasets/data.py: SOS, co-edge tokens with SEP between faces, EOS, then PAD):
    adds `label` [N, T] int64 and `label_mask` [N, T] bool (True = PAD) to a copy of `batch`.  Tokens address un-masked memory rows only."""
    rng = np.random.default_rng([seed, 15485863])
    T = cfg.label_seq_length
    n = batch["input"].shape[0]
    label = np.zeros((n, T), np.int64)
    nvalid = (~batch["input_mask"]).sum(1)
    for i in range(n):
        length = int(rng.integers(3, T - 1))
        seq = rng.integers(cfg.num_token, cfg.num_token + max(1, int(nvalid[i])), size=length)
        seq[rng.random(length) < 0.15] = 2                       # token.SEP
        label[i, 0] = 1                                          # token.SOS
        label[i, 1:1 + length] = seq
        label[i, 1 + length] = 3                                 # token.EOS
    out = dict(batch)
    out["label"] = label
    out["label_mask"] = label == 0
    return out


def quantize_state_dict(sd, bits: int = 8):
    """Snap every >= 2-d float tensor to a per-tensor symmetric integer grid (w = q * scale) so that a trained checkpoint
    can be committed as a small fixture.  Returns the npz payload: ``name`` (1-d / int tensors as they are) or
    ``name::q`` (int8) + ``name::scale`` (fp32 scalar)."""
    out = {}
    qmax = float(2 ** (bits - 1) - 1)
    for k, v in sd.items():
        v = np.asarray(v)
        if v.dtype.kind == "f" and v.ndim >= 2:
            scale = np.float32(np.abs(v).max() / qmax) if np.abs(v).max() > 0 else np.float32(1.0)
            out[k + "::q"] = np.clip(np.round(v / scale), -qmax, qmax).astype(np.int8)
            out[k + "::scale"] = scale
        else:
            out[k] = v
    return out


def load_state_dict_npz(path):
    """Inverse of np.savez(**sd) / np.savez(**quantize_state_dict(sd)): fp32 tensors by reference state_dict name."""
    with np.load(path) as z:
        raw = {k: z[k] for k in z.files}
    sd = {}
    for k, v in raw.items():
        if k.endswith("::scale"):
            continue
        if k.endswith("::q"):
            name = k[:-3]
            sd[name] = (v.astype(np.float32) * np.float32(raw[name + "::scale"])).astype(np.float32)
        else:
            sd[k] = v
    return sd
