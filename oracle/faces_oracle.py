"""CPU restatement of the reference's prediction parsing + enclosedness filter -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this module; the product path
(faceformer_b200.Engine.parse_faces -> ffb_parse_faces) never does.

Follows, relative to /root/reference:
  parse_predicts                 faceformer/trainer.py:196-206   (Trainer.parse_parallel_faces, predict half)
  e1_connects_e2                 dataset/tests/check_faces_enclosed.py:11-13
  is_face_enclosed               dataset/tests/check_faces_enclosed.py:18-46
  filter_faces_by_encloseness    faceformer/post_processing.py:8-20
Pinned by tests/golden/faces.npz: outputs of the reference's own functions (parse_parallel_faces executed from the source text
of trainer.py, filter_faces_by_encloseness imported) on seeded cases, oracle/make_golden_faces.py.
"""
import numpy as np

TOKEN_LEN, FACE_TYPE_OFFSET = 4, 1          # config.py:40-48


def parse_predicts(predicts, num_edges):
    faces = []
    for predict in np.array(predicts, dtype=np.int64):       # a copy: the reference subtracts in place
        cut = np.where((predict >= FACE_TYPE_OFFSET) & (predict < TOKEN_LEN))[0] + 1
        predict = np.split(predict, cut)[0]
        face_type = predict[-1] - FACE_TYPE_OFFSET
        predict = predict - TOKEN_LEN
        predict = predict[predict >= 0]
        predict = predict[predict < num_edges]
        if len(predict) > 0:
            faces.append((int(face_type), tuple(predict.tolist())))
    return faces


def e1_connects_e2(e1, e2, tol):
    return abs(e1[-1][0] - e2[0][0]) < tol and abs(e1[-1][1] - e2[0][1]) < tol


def is_face_enclosed(edges, face_indices, tol):
    all_loops, curr_loop, to_close, last_edge = [], [], None, None
    for ind in face_indices:
        if ind < len(edges):
            edge = edges[ind]
        else:
            continue
        if to_close is None:
            to_close = edge
        elif not e1_connects_e2(last_edge, edge, tol):
            return False
        last_edge = edge
        curr_loop.append(ind)
        if e1_connects_e2(edge, to_close, tol):
            to_close = None
            all_loops.append(curr_loop)
            curr_loop = []
    return all_loops if to_close is None else False


def filter_faces_by_encloseness(edges, faces, tol):
    out = []
    for face_type, face in faces:
        loops = is_face_enclosed(edges, face, tol)
        if loops:
            loops = [tuple(np.roll(loop, -np.argmin(loop), axis=0).astype(int).tolist()) for loop in loops]
            loops = sorted(loops, key=lambda x: x[0])
            out.append((face_type, tuple(loops)))
    return out


def synth_case(n, num_lines, T, seed):
    """Polygon wireframes (closed loops of straight edges, global order shuffled) + a predict tensor [n, F, T] that mixes true loops
    (every rotation), multi-loop faces, loops broken by one wrong edge, open chains, rows without a type token, rows that start
    with one, out-of-range indices and random rows.  Returns (wireframes: list of edge lists, predict int64)."""
    rng = np.random.default_rng([seed, 15485863])
    wfs, loops_all = [], []
    for _ in range(n):
        polys, total = [], 0
        for _ in range(int(rng.integers(2, 7))):
            k = int(rng.integers(3, 7))
            if total + k > num_lines:
                break
            c, r = rng.uniform(-0.6, 0.6, 2), rng.uniform(0.1, 0.4)
            ang = np.sort(rng.uniform(0, 2 * np.pi, k))
            polys.append(c[None] + r * np.stack([np.cos(ang), np.sin(ang)], 1))
            total += k
        perm = rng.permutation(total)
        edges, loops, g = [None] * total, [], 0
        for v in polys:
            loop = []
            for j in range(len(v)):
                a, b = v[j], v[(j + 1) % len(v)]
                if rng.random() < 0.15:                      # perturb an end point by ~tol: some joints sit on either side of it
                    b = b + rng.uniform(-3e-4, 3e-4, 2)
                mid = [(a + (b - a) * t).tolist() for t in np.linspace(0, 1, int(rng.integers(2, 6)))]
                mid[0], mid[-1] = a.tolist(), b.tolist()
                edges[int(perm[g])] = mid
                loop.append(int(perm[g])); g += 1
            loops.append(loop)
        wfs.append(edges)
        loops_all.append(loops)
    F = max(len(e) for e in wfs)
    pred = np.zeros((n, F, T), np.int64)
    for w in range(n):
        ne, loops = len(wfs[w]), loops_all[w]
        for f in range(F):
            kind = int(rng.integers(0, 9))
            row = rng.integers(0, ne + 8, T)                                 # garbage tail (may contain type tokens)
            loop = loops[int(rng.integers(len(loops)))]
            seq = np.roll(loop, int(rng.integers(len(loop)))).tolist()
            if kind == 1 and len(loops) > 1:                                 # two loops in one face
                other = loops[(loops.index(loop) + 1) % len(loops)]
                seq = seq + np.roll(other, int(rng.integers(len(other)))).tolist()
            elif kind == 2:                                                  # one wrong edge
                seq[int(rng.integers(len(seq)))] = int(rng.integers(ne))
            elif kind == 3:                                                  # open chain
                seq = seq[:-1]
            elif kind == 4:                                                  # out-of-range / special tokens interleaved
                seq = seq[:1] + [ne + 3] + seq[1:]
            body = (np.asarray(seq) + TOKEN_LEN).tolist()
            if kind == 5:
                row = rng.integers(TOKEN_LEN, ne + TOKEN_LEN, T)             # no type token at all
            elif kind == 6:
                row[0] = int(rng.integers(1, 4))                             # type token first
            elif kind == 7:
                pass                                                         # fully random row
            else:
                body = body[:T - 1]
                row[:len(body)] = body
                row[len(body)] = int(rng.integers(1, 4))
            pred[w, f] = row
    return wfs, pred
