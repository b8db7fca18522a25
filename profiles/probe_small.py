"""Small-batch regime probe (the reference's own test loop is batch = 1, trainer.py:51; BASELINE configs[0] is seq2seq N = 1):
device time per forward_eval and launches for N = 1 wireframes, per tensor-core mode.  Run on a GPU box:
    python profiles/probe_small.py [--json out.json]
"""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from faceformer_b200 import synth  # noqa: E402
from faceformer_b200.config import MODE_PARALLEL, MODE_SEQ2SEQ, OURS, SEQ2SEQ  # noqa: E402
from faceformer_b200.engine import Engine  # noqa: E402
from faceformer_b200.lib import FFB_OPT_PDL, FFB_OPT_TENSOR_CORE  # noqa: E402

OUT = []


def timed(eng, args, reps=3):
    eng.forward_eval(*args)
    torch.cuda.synchronize()
    best, steps = None, 0
    for _ in range(reps):
        l0 = eng.kernel_launches()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record()
        _, steps = eng.forward_eval(*args)
        e1.record()
        torch.cuda.synchronize()
        wall = (time.perf_counter() - t0) * 1e3
        ms = e0.elapsed_time(e1)
        if best is None or ms < best[0]:
            best = (ms, wall, eng.kernel_launches() - l0)
    return best + (steps,)


def main():
    for tc, pdl in ((0, 0), (1, 0), (1, 1)):
        # seq2seq, one 64-edge wireframe, 258 steps (the golden's weights: no early EOS)
        from util import load_case
        g = load_case("seq2seq_single64")
        e = Engine(g["cfg"], g["mode"], 0)
        e.load_state_dict(g["sd"])
        e.set_option(FFB_OPT_TENSOR_CORE, tc)
        e.set_option(FFB_OPT_PDL, pdl)
        b = g["batch"]
        args = (torch.from_numpy(b["input"]).cuda().flatten(2), torch.from_numpy(b["input_mask"]).cuda(), None)
        ms, wall, launches, steps = timed(e, args)
        rec = dict(workload="seq2seq_n1_64", tc=tc, pdl=pdl, ms=ms, wall_ms=wall, launches=launches, steps=steps, edges_per_s=steps / ms * 1e3,
                   us_per_step=ms / steps * 1e3)
        print(json.dumps(rec), flush=True)
        OUT.append(rec)
        e.close()
        # ours.yml, N = 1, several sizes
        sd = synth.synth_state_dict(OURS, MODE_PARALLEL, 0, "diverse")
        e = Engine(OURS, MODE_PARALLEL, 0)
        e.load_state_dict(sd)
        e.set_option(FFB_OPT_TENSOR_CORE, tc)
        e.set_option(FFB_OPT_PDL, pdl)
        for n in (24, 64, 120, 216):
            bt = synth.synth_batch(OURS, MODE_PARALLEL, 1, seed=n, num_edges=np.array([n], np.int64))
            args = (torch.from_numpy(bt["input"]).cuda().flatten(2), torch.from_numpy(bt["input_mask"]).cuda(), torch.from_numpy(bt["num_input"]).cuda())
            ms, wall, launches, steps = timed(e, args)
            rec = dict(workload=f"ours_n1_{n}", tc=tc, pdl=pdl, ms=ms, wall_ms=wall, launches=launches, steps=steps, edges_per_s=n * steps / ms * 1e3,
                       us_per_step=ms / steps * 1e3)
            print(json.dumps(rec), flush=True)
            OUT.append(rec)
        e.close()
    if "--json" in sys.argv:
        with open(sys.argv[sys.argv.index("--json") + 1], "w") as f:
            json.dump(OUT, f, indent=1)


if __name__ == "__main__":
    main()
