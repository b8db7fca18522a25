"""Distance of the CUDA path's pointer logits from the reference's fp32 result (d32), from the reference evaluated in float64
(d64), next to the reference's own fp32-vs-float64 noise (n32), per golden case and pipeline.  With tests/golden/noise_budget.npz
(oracle/make_golden_noise.py) also the per-stage budget: `mem` = max |our memory - float64 memory| next to the reference's own fp32
memory error, and `dec+head` = d64 of the last-step logits when the float64 memory is injected with ffb_set_memory (what is left is
the decoder stack + project + pointer dot); the reference-side split of n32 by stage is printed from the fixture.  Run on a GPU box:
    python profiles/logit_noise.py [--json out.json]
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from faceformer_b200.engine import Engine  # noqa: E402
from faceformer_b200.lib import FFB_OPT_DEDUP_PAD, FFB_OPT_TC_FORMAT, FFB_OPT_TENSOR_CORE  # noqa: E402
from util import CASES_ALL, load_case  # noqa: E402

FMIN = np.finfo(np.float32).min


def dist(a, b, ref32):
    ok = np.asarray(ref32) != FMIN          # un-masked entries (the float64 run fills with finfo(float64).min)
    return float(np.max(np.abs(np.asarray(a, np.float64)[ok] - np.asarray(b, np.float64)[ok])))


BUDGET, BMETA = {}, {}
try:
    import json
    with np.load(os.path.join(ROOT, "tests", "golden", "noise_budget.npz")) as z:
        BMETA = json.loads(str(z["meta"]))
        BUDGET = {k.split("::")[0]: z[k] for k in z.files if k.endswith("::memory64")}
    for k, v in BMETA.items():
        print(f"reference fp32 noise by stage, {k}: " + ", ".join(f"{a} {b:.2e}" for a, b in v.items()))
except Exception as ex:      # noqa: BLE001
    print("no noise_budget.npz:", ex)
RESULTS = []

print(f"{'case':26} {'pipeline':12} {'max|logit|':>10} {'n32':>9} | last: {'d32':>9} {'d64':>9} | prefix: {'d32':>9} {'d64':>9}")
for name in CASES_ALL:
    g = load_case(name)
    if "last_logits64" not in g:
        continue
    b = g["batch"]
    for label, tc, fmt, amma in (("all fp32 simt", 0, 2, 0), ("simt+3xTF32", 0, 2, 1), ("tc fp16x2", 2, 2, 2), ("tc bf16x3", 2, 3, 2)):
        e = Engine(g["cfg"], g["mode"], 0)
        e.load_state_dict(g["sd"])
        try:
            e.set_option(FFB_OPT_TC_FORMAT, fmt)
            e.set_option(FFB_OPT_TENSOR_CORE, tc)
            e.set_option(6, amma)                                # FFB_OPT_ATTN_MMA: 0 = fp32 SIMT attention kernels
        except Exception:
            e.close()
            continue
        coords = torch.from_numpy(b["input"]).cuda().flatten(2)
        mask, ni = torch.from_numpy(b["input_mask"]).cuda(), torch.from_numpy(b["num_input"]).cuda()
        pred, steps = e.forward_eval(coords, mask, ni)
        same = bool(np.array_equal(pred.cpu().numpy(), g["predict"]))
        last = e.get_last_logits().cpu().numpy()
        e.set_option(FFB_OPT_DEDUP_PAD, 0)
        e.encode(coords, mask, ni)
        pre = e.forced_prefix_logits(torch.from_numpy(g["prefix"]).cuda()).cpu().numpy()
        mx = float(np.max(np.abs(g["last_logits64"][g["last_logits"] != FMIN])))
        L, P = g["last_logits"], g["prefix_logits"]
        n32 = max(dist(L, g["last_logits64"], L), dist(P, g["prefix_logits64"], P))
        print(f"{name:26} {label:12} {mx:10.2f} {n32:9.2e} | last: {dist(last, L, L):9.2e} {dist(last, g['last_logits64'], L):9.2e} | "
              f"prefix: {dist(pre, P, P):9.2e} {dist(pre, g['prefix_logits64'], P):9.2e}  tokens {'==' if same else '!='}", flush=True)
        rec = dict(case=name, pipeline=label, max_logit=mx, n32=n32, last_d32=dist(last, L, L), last_d64=dist(last, g["last_logits64"], L),
                   prefix_d32=dist(pre, P, P), prefix_d64=dist(pre, g["prefix_logits64"], P), tokens_equal=same)
        if name in BUDGET:
            T = g["cfg"].seq_len(g["mode"])
            m64 = BUDGET[name]
            ours = e.get_memory().cpu().numpy()
            vm = np.abs(m64).sum(-1) > 0
            rec["mem_err"] = float(np.max(np.abs(ours[vm].astype(np.float64) - m64[vm])))
            rec["mem_err_ref32"] = float(np.max(np.abs(g["memory"][vm].astype(np.float64) - m64[vm])))
            e.set_memory(torch.from_numpy(m64).cuda())
            pre_last = np.ascontiguousarray(g["predict"].reshape(-1, T)[:, :g["steps"]].T)
            inj = e.forced_prefix_logits(torch.from_numpy(pre_last).cuda()).cpu().numpy()
            rec["dec_head_d64"] = dist(inj, g["last_logits64"], L)
            print(f"{'':26} {'':12} budget: mem {rec['mem_err']:.2e} (reference fp32 {rec['mem_err_ref32']:.2e})  dec+head d64 {rec['dec_head_d64']:.2e} "
                  f"(reference fp32: decoder {BMETA[name]['decoder_fp32']:.2e}, head {BMETA[name]['head_fp32']:.2e})", flush=True)
        RESULTS.append(rec)
        e.close()
if "--json" in sys.argv:
    import json
    with open(sys.argv[sys.argv.index("--json") + 1], "w") as f:
        json.dump({"reference_budget": BMETA, "results": RESULTS}, f, indent=1)
