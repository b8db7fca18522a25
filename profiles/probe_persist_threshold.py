"""Where the persistent decode kernel beats the multi-kernel path: one ours.yml wireframe of n edges (B = n sequences, 36 steps) and N seq2seq
wireframes of 64 edges (B = N, 258 steps), FFB_OPT_PERSISTENT = 0 vs 2.  Prints ms per forward_eval (median of 5 after 2 warm-ups)."""
import os, sys, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, os.path.join(sys.path[0], "tests"))
import numpy as np, torch
from faceformer_b200 import synth
from faceformer_b200.config import MODE_PARALLEL, MODE_SEQ2SEQ, OURS, SEQ2SEQ
from faceformer_b200.engine import Engine
from faceformer_b200.lib import FFB_OPT_PERSISTENT

def run(cfg, mode, n_wf, n_edges, persist):
    sd = synth.synth_state_dict(cfg, mode, 0, "diverse")
    batch = synth.synth_batch(cfg, mode, n_wf, 7, lo=n_edges, hi=n_edges)
    e = Engine(cfg, mode, 0); e.load_state_dict(sd); e.set_option(FFB_OPT_PERSISTENT, persist)
    c = torch.from_numpy(batch["input"]).cuda().flatten(2); m = torch.from_numpy(batch["input_mask"]).cuda(); ni = torch.from_numpy(batch["num_input"]).cuda()
    ts = []
    for i in range(7):
        torch.cuda.synchronize(); t = time.perf_counter()
        pred, steps = e.forward_eval(c, m, ni)
        torch.cuda.synchronize(); ts.append((time.perf_counter() - t) * 1e3)
    used = e.used_persistent(); e.close()
    return float(np.median(ts[2:])), steps, used, pred.cpu().numpy()

out = []
for n in (8, 16, 24, 28, 40, 64):
    a, s, ua, pa = run(OURS, MODE_PARALLEL, 1, n, 0); b, _, ub, pb = run(OURS, MODE_PARALLEL, 1, n, 2)
    out.append(dict(workload=f"ours.yml 1 wireframe x {n} edges", rows=n * 36, steps=s, multi_kernel_ms=round(a, 2), persistent_ms=round(b, 2), used=[ua, ub], same_tokens=bool((pa == pb).all())))
    print(out[-1], flush=True)
for nw in (1, 2, 3, 4):
    a, s, ua, pa = run(SEQ2SEQ, MODE_SEQ2SEQ, nw, 64, 0); b, _, ub, pb = run(SEQ2SEQ, MODE_SEQ2SEQ, nw, 64, 2)
    out.append(dict(workload=f"seq2seq.yml {nw} wireframes x 64 edges", rows=nw * 258, steps=s, multi_kernel_ms=round(a, 2), persistent_ms=round(b, 2), used=[ua, ub], same_tokens=bool((pa == pb).all())))
    print(out[-1], flush=True)
json.dump(out, open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "probe_persist_threshold_r2.json"), "w"), indent=1)
