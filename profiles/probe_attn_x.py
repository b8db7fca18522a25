"""Run the tcgen05 attention kernel alone on bench-sized problems (for ncu captures and event timing).

    python profiles/probe_attn_x.py            # event-timed, prints us per launch
    ncu --set full --import-source on -k regex:attn_x -c 2 -o gpurun_out/attn_x python profiles/probe_attn_x.py --once
"""
import sys, time
import numpy as np, torch
sys.path.insert(0, ".")
from faceformer_b200.config import OURS, MODE_PARALLEL
from faceformer_b200.engine import Engine

once = "--once" in sys.argv
H = OURS.num_head
eng = Engine(OURS, MODE_PARALLEL, 0)
rng = np.random.default_rng(0)
cases = [("cross", 5, 32, 116 * 36, 150), ("self", 6, 3719, 36, 36), ("cross_mma_sync", 4, 32, 116 * 36, 150)]
for name, kind, G, nq, nk in cases:
    q = torch.from_numpy((rng.normal(size=(G * nq, H * 64)) * 1.5).astype(np.float32)).cuda()
    k = torch.from_numpy((rng.normal(size=(G * nk, H * 64)) * 1.5).astype(np.float32)).cuda()
    v = torch.from_numpy((rng.normal(size=(G * nk, H * 64)) * 1.5).astype(np.float32)).cuda()
    reps = 1 if once else 5
    best = 1e9
    for _ in range(reps):
        torch.cuda.synchronize(); eng.set_option(4, 1)           # FFB_OPT_PROFILE: event pair around every launch
        eng.op_attention(kind, q, k, v, G, nq, nk)
        prof = eng.profile_read(); eng.set_option(4, 0)
        ms = prof["attn_tiled"]["ms"]
        best = min(best, ms)
    items = (G * ((nq + 127) // 128) if kind != 6 else (G + (128 // nq) - 1) // (128 // nq)) * H
    print(f"{name}: G={G} nq={nq} nk={nk} items={items} kernel {best*1e3:.1f} us  -> {best*1e3/ (items/148):.2f} us per item per SM", flush=True)
