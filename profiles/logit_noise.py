"""Distance of the CUDA path's pointer logits from the reference's fp32 result (d32), from the reference evaluated in float64
(d64), next to the reference's own fp32-vs-float64 noise (n32), per golden case and pipeline.  Run on a GPU box:
    python profiles/logit_noise.py
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
        e.close()
