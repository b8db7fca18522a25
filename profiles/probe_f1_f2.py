"""Throughput of the steps right before / after the path at the bench geometry: 32 wireframes, ours.yml
(num_lines 216, T 37), against the reference's own per-edge / per-face Python (restated in oracle/, pinned to the reference's outputs).

  f1  ffb_featurize   raw edge polylines -> input [N, 216, 50, 2] f32, input_mask, num_input   (datasets/data_para.py:8-25,59-68)
  f2  ffb_parse_faces predict [N, F, 37] -> faces (type, loops) + enclosedness filter          (trainer.py:196-206, post_processing.py:8-20)

GPU time = CUDA events around the C-ABI call with device-resident inputs (median of 20); "api" = the Python method incl. ragged flattening of
the nested lists, H2D of the points and building the nested result lists; "api_flat" = the same with the wireframes flattened once beforehand
(Engine.flatten_wireframes).  CPU = the oracle restatement on one core (what a DataLoader worker /
face_accuracy runs per sample)."""
import ctypes as C, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from faceformer_b200.config import MODE_PARALLEL, OURS
from faceformer_b200.engine import Engine, _ptr
from faceformer_b200.lib import FFB_DEVICE
from oracle import featurize_oracle as fe, faces_oracle as fa

N, NL, T = 32, 216, 37
e = Engine(OURS, MODE_PARALLEL, 0)
dev = torch.device("cuda", 0)

def med_ms(fn, k=20):
    ts = []
    for _ in range(k + 3):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    return float(np.median(ts[3:]))

out = {}
# ---- f1
wfs = fe.synth_wireframes(N, NL, 5, lo=24)
n_edges = sum(len(w) for w in wfs)
pts, eoff, woff = e._ragged(wfs)
t_pts, t_e, t_w = (torch.from_numpy(x).to(dev) for x in (pts, eoff, woff))
o = torch.empty((N, NL, 50, 2), dtype=torch.float32, device=dev); m = torch.empty((N, NL), dtype=torch.uint8, device=dev); ni = torch.empty((N,), dtype=torch.int64, device=dev)
call = lambda: e._check(e._lib.ffb_featurize(e._h, _ptr(t_pts), _ptr(t_e), _ptr(t_w), N, _ptr(o), _ptr(m), _ptr(ni), FFB_DEVICE, e._stream()))
gpu_ms = med_ms(call)
t0 = time.perf_counter(); e.featurize(wfs); torch.cuda.synchronize(); api_ms = (time.perf_counter() - t0) * 1e3
flat = Engine.flatten_wireframes(wfs)
e.featurize(flat); torch.cuda.synchronize()
t0 = time.perf_counter(); e.featurize(flat); torch.cuda.synchronize(); api_flat_ms = (time.perf_counter() - t0) * 1e3
t0 = time.perf_counter(); ref = fe.featurize(wfs, NL); cpu_ms = (time.perf_counter() - t0) * 1e3
assert np.array_equal(o.cpu().numpy().view(np.uint32), ref[0].view(np.uint32))
bytes_moved = pts.nbytes + o.numel() * 4
out["f1_featurize"] = dict(wireframes=N, edges=n_edges, points=int(pts.shape[0]), gpu_ms=round(gpu_ms, 4), api_ms=round(api_ms, 2), api_flat_ms=round(api_flat_ms, 3), cpu_oracle_ms=round(cpu_ms, 1),
                           gpu_edges_per_s=round(n_edges / gpu_ms * 1e3), cpu_edges_per_s=round(n_edges / cpu_ms * 1e3), gpu_gb_per_s=round(bytes_moved / gpu_ms / 1e6, 1),
                           bit_exact_vs_oracle=True)
# ---- f2
wfs2, pred = fa.synth_case(N, NL, T, 7)
F = pred.shape[1]
p_dev = torch.from_numpy(pred).to(dev)
pts2, eoff2, woff2 = e._ragged(wfs2)
ins = [torch.from_numpy(x).to(dev) for x in (pts2, eoff2, woff2)]
valid = torch.empty((N, F), dtype=torch.uint8, device=dev)
ft, nl_, ni2 = (torch.empty((N, F), dtype=torch.int32, device=dev) for _ in range(3))
ll, idx = (torch.empty((N, F, T), dtype=torch.int32, device=dev) for _ in range(2))
call2 = lambda: e._check(e._lib.ffb_parse_faces(e._h, _ptr(p_dev), N, F, _ptr(ins[0]), _ptr(ins[1]), _ptr(ins[2]), 2e-4, 1, _ptr(valid), _ptr(ft), _ptr(nl_), _ptr(ll),
                                                 _ptr(idx), _ptr(ni2), FFB_DEVICE, e._stream()))
gpu2 = med_ms(call2)
t0 = time.perf_counter(); got = e.parse_faces(p_dev, wfs2, tol=2e-4); api2 = (time.perf_counter() - t0) * 1e3
flat2 = Engine.flatten_wireframes(wfs2)
t0 = time.perf_counter(); got_flat = e.parse_faces(p_dev, flat2, tol=2e-4); api2_flat = (time.perf_counter() - t0) * 1e3
assert got_flat == got
t0 = time.perf_counter(); want = [fa.filter_faces_by_encloseness(wfs2[w], fa.parse_predicts(pred[w], len(wfs2[w])), 2e-4) for w in range(N)]; cpu2 = (time.perf_counter() - t0) * 1e3
assert got == want
out["f2_parse_faces"] = dict(wireframes=N, sequences=int(N * F), faces_kept=int(sum(len(f) for f in got)), gpu_ms=round(gpu2, 4), api_ms=round(api2, 2), api_flat_ms=round(api2_flat, 3), cpu_oracle_ms=round(cpu2, 1),
                             gpu_sequences_per_s=round(N * F / gpu2 * 1e3), cpu_sequences_per_s=round(N * F / cpu2 * 1e3), equal_to_oracle=True)
print(json.dumps(out, indent=1))
json.dump(out, open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "probe_f1_f2_r2.json"), "w"), indent=1)
e.close()
