"""Encoder-only throughput on synthetic 2048-edge wireframes (BASELINE.json configs[4] geometry: num_lines 2048, L = 2052).
    python profiles/probe_encoder.py [n_wireframes]
Prints wireframes/s and TFLOP/s (104.7 GFLOP per wireframe: 51.6 linear + 51.7 attention + 1.3 embedding) for the
tensor-core encoder (fp16x2 tcgen05 GEMMs + fp16x2 mma.sync attention for > 256 keys) and the fp32 SIMT encoder."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from faceformer_b200 import synth
from faceformer_b200.config import MODE_PARALLEL, OURS
from faceformer_b200.engine import Engine
from faceformer_b200.lib import FFB_OPT_ENCODER_TC

n = int(sys.argv[1]) if len(sys.argv) > 1 else 16
cfg = OURS.replace(num_lines=2048)
sd = synth.synth_state_dict(cfg, MODE_PARALLEL, 6, "diverse")
batch = synth.synth_batch(cfg, MODE_PARALLEL, n, seed=8, num_edges=np.full(n, 2048, np.int64))
coords = torch.from_numpy(batch["input"]).cuda().flatten(2)
mask, ni = torch.from_numpy(batch["input_mask"]).cuda(), torch.from_numpy(batch["num_input"]).cuda()
mems = []
for enc_tc in (1, 0):
    e = Engine(cfg, MODE_PARALLEL, 0)
    e.load_state_dict(sd)
    e.set_option(FFB_OPT_ENCODER_TC, enc_tc)
    e.encode(coords, mask, ni); torch.cuda.synchronize()
    t = []
    for _ in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); e.encode(coords, mask, ni); e1.record(); torch.cuda.synchronize()
        t.append(e0.elapsed_time(e1))
    ms = min(t)
    mems.append(e.get_memory().cpu().numpy())
    print(f"encoder_tc={enc_tc}: {n} wireframes x 2048 edges: {ms:.1f} ms -> {n / ms * 1e3:.1f} wireframes/s, {n * 104.7 / ms:.1f} TFLOP/s (incl. the cross K/V cache), fallbacks {e.fp16_fallbacks()}", flush=True)
    e.close()
print("max |memory_tc - memory_simt| =", float(np.max(np.abs(mems[0] - mems[1]))))
