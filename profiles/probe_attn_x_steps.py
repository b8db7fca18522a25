import sys
import numpy as np, torch
sys.path.insert(0, ".")
from faceformer_b200.config import OURS, MODE_PARALLEL
from faceformer_b200.engine import Engine
H = OURS.num_head
eng = Engine(OURS, MODE_PARALLEL, 0)
rng = np.random.default_rng(0)
for name, kind, G, nq, nk in [("cross P=1", 5, 32, 116, 150), ("cross P=4", 5, 32, 116*4, 150), ("cross P=12", 5, 32, 116*12, 150), ("cross P=36", 5, 32, 116*36, 150),
                              ("self P=1", 6, 3719, 1, 1), ("self P=4", 6, 3719, 4, 4), ("self P=12", 6, 3719, 12, 12), ("self P=24", 6, 3719, 24, 24), ("self P=32", 6, 3719, 32, 32), ("self P=36", 6, 3719, 36, 36)]:
    q = torch.from_numpy((rng.normal(size=(G * nq, H * 64)) * 1.5).astype(np.float32)).cuda()
    k = torch.from_numpy((rng.normal(size=(G * nk, H * 64)) * 1.5).astype(np.float32)).cuda()
    v = torch.from_numpy((rng.normal(size=(G * nk, H * 64)) * 1.5).astype(np.float32)).cuda()
    best = 1e9
    for _ in range(4):
        torch.cuda.synchronize(); eng.set_option(4, 1)
        eng.op_attention(kind, q, k, v, G, nq, nk)
        prof = eng.profile_read(); eng.set_option(4, 0)
        best = min(best, prof["attn_tiled"]["ms"])
    print(f"{name}: kernel {best*1e3:.1f} us", flush=True)
