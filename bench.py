#!/usr/bin/env python
"""bench.py -- decoded edges/sec (greedy) of the FaceFormer pointer-decode path on B200.

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...   # the CPU arm (unmodified torch reference, oracle/_ref), rank 0 only

A "step" is one pass of the hot path (model(batch): embedding + encoder + cross-K/V + the whole
greedy loop) over one batch of synthetic wireframes.  Workload (BASELINE.json configs[1]):
configs/ours.yml geometry (E=512, H=8, FF=1024, 6+6 layers, num_lines=216, T=37), batch=32
wireframes per GPU, n_edges ~ U[24,216], greedy decode.  Metric: decoded edges/s = B_eff*S / time, with
B_eff the DISTINCT sequences of the batch (the reference decodes B = N*F slots, of which the F - n_i padded
anchors of a wireframe are copies of one sequence; `value_slots` = B*S / time is the SURVEY.md 8d count) and S the
executed decode steps.  Both arms report both counts.

Keys beyond the base contract:
  value     device-resident inputs/outputs, CUDA-event timed on the launching stream, max over ranks
  e2e       same metric through the C ABI with HOST buffers (pinned): H2D of inputs and D2H of
            `predict` inside the timed region
  roofline  the dominant kernel (tc::gemm_kernel): algorithmic FLOPs per launch / event-timed duration,
            measured on one extra (untimed) step with an event pair around every launch
  cpu_baseline  the unmodified torch reference (oracle/_ref) timed on this box's host cores on a bounded stratified sample
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from faceformer_b200 import synth  # noqa: E402
from faceformer_b200.config import MODE_PARALLEL, OURS  # noqa: E402

METRIC = "decoded_edges_per_sec"
UNIT = "edges/s"
WORKLOAD = "configs/ours.yml greedy decode, batch=32 wireframes/GPU, num_lines=216, T=37, n_edges~U[24,216]"
COUNT_NOTE = ("value counts DISTINCT decoded sequences x steps (B_eff*S: padded-anchor slots are copies of one sequence and are not "
              "counted); value_slots counts the reference's output slots (N*F*S, SURVEY.md 8d)")


def load_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            p = json.load(f)
        return dict(hbm_gbs=float(p["hbm_gbs"]), bf16_burst=float(p["bf16_tflops"]),
                    bf16_sustained=float(p.get("bf16_tflops_sustained", p["bf16_tflops"])), source="measured")
    except Exception:
        return dict(hbm_gbs=6650.0, bf16_burst=1590.0, bf16_sustained=1400.0, source="fallback")


class ClockSampler(threading.Thread):
    """Samples nvidia-smi clocks / throttle reasons every 200 ms while the timed region runs."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self._halt = index, [], threading.Event()

    def run(self):
        while not self._halt.is_set():
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([x.strip() for x in out.split(",")])
            except Exception:
                pass
            self._halt.wait(0.2)

    def stop(self):
        self._halt.set()
        self.join(timeout=6)
        sm = [float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in self.rows if len(r) >= 7 for i in range(4) if r[3 + i].lower().startswith("active")})
        return dict(sm_mhz=float(np.median(sm)) if sm else None, sm_max_mhz=max(mx) if mx else None,
                    reasons=reasons, samples=len(self.rows))


# ---- CPU arm: the UNMODIFIED torch reference (oracle/_ref, copied by oracle/build_ref.py) on the host cores -------------------
# A bounded, stratified sample of the workload (the reference needs 73 GFLOP per decoded sequence): each entry is one
# model(batch) call.  "ragged" batches have padded anchors (F > n_i), like the N=32 GPU batch; singles are the reference's own test
# batch size (trainer.py:51).
CPU_SAMPLES = {
    "step": [[24, 30]],                      # one ragged N=2 batch per `--impl reference` step (~5 s on 16 cores)
    "baseline": [[24], [24, 30], [96]],      # the cpu_baseline leg of the default run (~20-40 s)
}


def use_all_cores():
    """torchrun exports OMP_NUM_THREADS=1; the CPU arm must still use every host core."""
    n = os.cpu_count() or 1
    try:
        import torch
        torch.set_num_threads(n)
    except Exception:
        pass
    try:
        from threadpoolctl import threadpool_limits
        threadpool_limits(limits=n)
    except Exception:
        pass
    return n


def load_traffic(kernel_tag):
    """DRAM bytes per launch of the dominant kernel from the committed ncu --set full capture (profiles/), or None."""
    try:
        with open(os.path.join(ROOT, "profiles", "ncu_traffic.json")) as f:
            ent = json.load(f).get(kernel_tag)
        return (ent["dram_bytes_per_launch"], ent.get("note")) if ent else (None, None)
    except Exception:
        return (None, None)


class CpuReference:
    """model(batch) of the unmodified reference (kind "reference"); falls back to the numpy oracle port (kind "port") only when
    oracle/_ref is absent."""

    def __init__(self, sd):
        self.cores = use_all_cores()
        self.sd = sd
        try:
            import torch
            sys.path.insert(0, os.path.join(ROOT, "oracle"))
            from build_ref import import_reference
            _, cls = import_reference()
            m = cls(**OURS.model_kwargs(MODE_PARALLEL)).eval()
            m.load_state_dict({k: torch.from_numpy(np.asarray(v)) for k, v in sd.items()}, strict=True)
            self.kind, self.model, self.torch = "reference", m, torch
        except Exception as e:                          # noqa: BLE001
            self.kind, self.model, self.why = "port", None, repr(e)

    def run(self, sample, seed):
        """-> dict(seconds, slots, distinct, steps): slots = sum N*F*S (what the reference decodes), distinct = the sequences that
        differ (padded anchors are copies of the row-3 sequence)."""
        sec = slots = distinct = 0
        steps = []
        for ne in sample:
            ne = np.asarray(ne, np.int64)
            batch = synth.synth_batch(OURS, MODE_PARALLEL, len(ne), seed=seed, num_edges=ne)
            t0 = time.perf_counter()
            if self.kind == "reference":
                with self.torch.no_grad():
                    pred = self.model({k: self.torch.from_numpy(v) for k, v in batch.items()})["predict"].numpy()
                flat = pred.reshape(-1, pred.shape[-1])
                S = next((s for s in range(1, flat.shape[1]) if np.all(flat[:, s] < 4)), flat.shape[1] - 1)
            else:
                from oracle import faceformer_oracle as orc
                S = orc.forward_eval(self.sd, OURS.to_dict(), MODE_PARALLEL, batch)["steps"]
            sec += time.perf_counter() - t0
            F = int(ne.max())
            slots += len(ne) * F * S
            distinct += int(sum(n + (1 if (n < F and n < 4) else 0) for n in ne)) * S
            steps.append(int(S))
        return dict(seconds=sec, slots=slots, distinct=distinct, steps=steps)

    def describe(self, sample, r):
        what = ("unmodified torch reference (oracle/_ref: faceformer/models/model_para.py forward_eval), fp32, torch "
                f"{self.torch.__version__}" if self.kind == "reference" else "numpy port (oracle/faceformer_oracle.py; oracle/_ref missing)")
        return (f"{what}; model(batch) calls with n_edges {sample} (ragged batches decode N*F slots incl. padded anchors), full ours.yml "
                f"model, steps {r['steps']}: {r['distinct']} distinct / {r['slots']} slot edges in {r['seconds']:.1f} s")


def run_reference(args):
    """CPU arm: the reference's own CPU implementation on all host cores, one bounded sample per step."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    sd = synth.synth_state_dict(OURS, MODE_PARALLEL, args.seed, "diverse")
    ref = CpuReference(sd)
    sample = CPU_SAMPLES["step"]
    times, r = [], None
    for i in range(args.warmup + args.steps):
        r = ref.run(sample, args.seed)
        if i >= args.warmup:
            times.append(r["seconds"])
    dt = float(np.mean(times))
    value, value_slots = r["distinct"] / dt, r["slots"] / dt
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "sample": f"one model(batch) call per step, n_edges {sample}", "edge_count": COUNT_NOTE},
        "value_slots": value_slots,
        "cpu_baseline": {"value": value, "value_slots": value_slots, "unit": UNIT, "cores": ref.cores, "kind": ref.kind,
                         "sample": ref.describe(sample, r)},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0}))


def bench_batch_parity(eng, pred_d, batch, cfg):
    """The timed batch IS the golden-pinned one (tests/golden/bench_batch.npz: the unmodified reference's output for seed 0, 32 wireframes):
    compare the tokens of the last timed step and the plain max-abs distance of the last-step logits (two wireframes) with the reference."""
    try:
        with np.load(os.path.join(ROOT, "tests", "golden", "bench_batch.npz")) as z:
            g = {k: z[k] for k in z.files}
        pred = pred_d.cpu().numpy()
        want = g["predict"].astype(np.int64)
        F = int(batch["num_input"].max())
        lg = eng.get_last_logits().cpu().numpy().reshape(len(batch["num_input"]), F, cfg.mem_len)
        fmin = np.finfo(np.float32).min
        out = {"fixture": "tests/golden/bench_batch.npz (reference output for exactly this batch)", "tokens_equal": bool(np.array_equal(pred, want)),
               "token_mismatches": int((pred != want).sum()), "logits": {}}
        for tag, w in zip("ab", g["logit_wireframes"]):
            n = int(batch["num_input"][w])
            ref32, ref64, ours = g[f"last_logits_{tag}"], g[f"last_logits64_{tag}"], lg[w, :n]
            ok = ref32 != fmin
            out["logits"][f"wireframe_{int(w)}"] = {
                "max_abs_diff_vs_reference_fp32": float(np.max(np.abs(ours[ok].astype(np.float64) - ref32[ok]))),
                "max_abs_diff_vs_reference_float64": float(np.max(np.abs(ours[ok].astype(np.float64) - ref64[ok]))),
                "reference_fp32_vs_float64": float(np.max(np.abs(ref32[ok].astype(np.float64) - ref64[ok]))),
                "max_abs_logit": float(np.max(np.abs(ref32[ok]))), "masked_pattern_equal": bool(np.array_equal(ours == fmin, ref32 == fmin))}
        return out
    except Exception as e:                           # noqa: BLE001
        return {"error": repr(e)}


def run_ours(args):
    import torch
    import torch.distributed as dist
    from faceformer_b200.engine import Engine, pack_state_dict
    from faceformer_b200.lib import FFB_OPT_PROFILE, FFB_OPT_TC_FORMAT, FFB_OPT_TENSOR_CORE
    from faceformer_b200 import sharding

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a GPU: faceformer_b200 has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    cfg, mode = OURS, MODE_PARALLEL
    T = cfg.max_face_length

    # weights: rank 0 builds the blob, NCCL broadcast over NVLink
    eng = Engine(cfg, mode, local)
    nw = eng.weight_count()
    if rank == 0:
        sd = synth.synth_state_dict(cfg, mode, args.seed, "diverse")
        blob = torch.from_numpy(pack_state_dict(sd, cfg, mode)).to(dev)
    else:
        sd, blob = None, torch.empty(nw, dtype=torch.float32, device=dev)
    sharding.broadcast_weights(blob, 0)
    eng.load_blob(blob)
    eng.set_option(FFB_OPT_TC_FORMAT, args.tc_format)
    eng.set_option(FFB_OPT_TENSOR_CORE, args.tc)
    if args.pdl >= 0:
        from faceformer_b200.lib import FFB_OPT_PDL
        eng.set_option(FFB_OPT_PDL, args.pdl)
    if args.gemm_variant >= 0:
        from faceformer_b200.lib import FFB_OPT_GEMM_VARIANT
        eng.set_option(FFB_OPT_GEMM_VARIANT, args.gemm_variant)
    if args.attn_x >= 0:
        from faceformer_b200.lib import FFB_OPT_ATTN_X
        eng.set_option(FFB_OPT_ATTN_X, args.attn_x)

    # one batch of 32 wireframes per rank (weak scaling: the number of wireframes per GPU is fixed).  Rank 0 decodes the golden-pinned
    # batch (seed); every other rank draws its own batch (seed + rank): edge counts, F and the sequence count differ per rank, so the
    # max-over-ranks time carries the load imbalance a real sharded test split has.
    batch = synth.synth_batch(cfg, mode, args.batch, seed=args.seed + rank)
    N = args.batch
    coords_h = torch.from_numpy(batch["input"].reshape(N, cfg.num_lines, -1)).pin_memory()
    mask_h = torch.from_numpy(batch["input_mask"].astype(np.uint8)).pin_memory()
    ni_h = torch.from_numpy(batch["num_input"]).pin_memory()
    coords_d, mask_d, ni_d = coords_h.to(dev), mask_h.to(dev), ni_h.to(dev)
    F = int(batch["num_input"].max())
    pred_d = torch.empty((N, F, T), dtype=torch.int64, device=dev)
    pred_h = torch.empty((N, F, T), dtype=torch.int64).pin_memory()
    gather_in = torch.full((N, cfg.num_lines, T), -1, dtype=torch.int32, device=dev)     # fixed shape: F differs per rank
    gather_out = torch.empty((world, N, cfg.num_lines, T), dtype=torch.int32, device=dev) if world > 1 else None

    def gather():
        if world > 1:       # all-gather of the predicted face loops (fixed shape, int32)
            gather_in[:, :F] = pred_d.to(torch.int32)
            dist.all_gather_into_tensor(gather_out, gather_in)

    def step_device():
        _, s = eng.forward_eval(coords_d, mask_d, ni_d, want_steps=True, out=pred_d)
        gather()
        return s

    def step_host():
        _, s = eng.forward_eval(coords_h.numpy(), mask_h.numpy(), ni_h.numpy(), want_steps=True, out=pred_h.numpy())
        if world > 1:
            pred_d.copy_(pred_h, non_blocking=True)
            gather()
        return s

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, k):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        s = 0
        for _ in range(k):
            s = fn()
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item()), s

    for _ in range(max(1, args.warmup)):          # at least one pass so that the batch geometry is known
        S = step_device()
    info = eng.batch_info()
    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start()
    l0 = eng.kernel_launches()
    ms, S = timed(step_device, args.steps)
    launches = eng.kernel_launches() - l0
    clocks = sampler.stop() if sampler else None

    # [distinct sequences x steps (the work done), reference output slots x steps]
    edges_local = torch.tensor([float(info["B_eff"] * S), float(info["B"] * S)], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(edges_local)
    edges_per_step, slots_per_step = float(edges_local[0].item()), float(edges_local[1].item())
    value = edges_per_step * args.steps / (ms / 1e3)
    value_slots = slots_per_step * args.steps / (ms / 1e3)

    # end to end through the C ABI with host buffers
    step_host()
    e2e_steps = max(1, min(args.steps, 3))
    ms_e, _ = timed(step_host, e2e_steps)
    e2e_value = edges_per_step * e2e_steps / (ms_e / 1e3)
    h2d = coords_h.numel() * 4 + mask_h.numel() + ni_h.numel() * 8
    d2h = pred_h.numel() * 8

    # per-kernel-class breakdown and roofline of the dominant kernel, on one extra untimed step
    eng.set_option(FFB_OPT_PROFILE, 1)
    step_device()
    prof = eng.profile_read()
    eng.set_option(FFB_OPT_PROFILE, 0)
    if rank == 0:
        peaks = load_peaks()
        tot_ms = sum(v["ms"] for v in prof.values())
        use_tc = prof["linear_tc"]["ms"] > prof["linear"]["ms"]
        lin = prof["linear_tc"] if use_tc else prof["linear"]
        ach = lin["flops"] / (lin["ms"] * 1e-3) / 1e12 if lin["ms"] > 0 else 0.0
        # MMA passes of the precision mode: fp16x2 split = 3 fp16 MMAs per fp32-equivalent product, bf16x3 = 6;
        # the fp32 SIMT kernel does not use the tensor pipe at all and is compared with the un-divided peak.
        fmt = 3 if eng.fp16_fallbacks() else args.tc_format
        passes = (3 if fmt == 2 else 6) if use_tc else 1
        peak = peaks["bf16_sustained"] / passes
        traffic, traffic_note = load_traffic(f"gemm_kernel<{fmt}>" if use_tc else "linear_kernel")
        kname = (f"tc::gemm_kernel<{fmt}> (tcgen05 {'fp16x2' if fmt == 2 else 'bf16x3'} split, {passes} MMA passes)" if use_tc
                 else "linear_kernel (fp32 SIMT FFMA)")
        roofline = {"bound": "tensor", "kernel": kname, "achieved": ach, "peak": peak, "unit": "TFLOP/s", "mma_passes": passes,
                    "frac": ach / peak, "traffic": traffic, "traffic_note": traffic_note, "peak_source": f"MEASURED_PEAKS.json bf16_tflops_sustained / {passes} ({peaks['source']})",
                    "avg_launch_ms": lin["ms"] / max(1, lin["launches"]), "launches_per_step": lin["launches"],
                    "share_of_step": lin["ms"] / tot_ms if tot_ms else None,
                    "fp32_ffma_peak_tflops": 148 * 128 * 2 * (clocks["sm_mhz"] or 1965.0) * 1e6 / 1e12 if clocks else None,
                    # whole-step view against the HBM roofline: algorithmic bytes of the decode (66 KB per (row, layer) for the
                    # Ld-1 full decoder layers -- GEMM operands / outputs 40, LayerNorm 14, self-attention 8, cross-attention 4 KB -- and 18 KB
                    # for the pruned last layer), rows = B_eff * S(S+1)/2, over the device-timed step
                    "step_hbm": (lambda rows, t: {"algorithmic_bytes": rows * ((cfg.num_decoder_layers - 1) * 66e3 + 18e3),
                                                  "achieved_gbs": rows * ((cfg.num_decoder_layers - 1) * 66e3 + 18e3) / t / 1e9,
                                                  "peak_gbs": peaks["hbm_gbs"],
                                                  "frac": rows * ((cfg.num_decoder_layers - 1) * 66e3 + 18e3) / t / 1e9 / peaks["hbm_gbs"]})(
                        info["B_eff"] * S * (S + 1) / 2.0, ms / args.steps / 1e3),
                    "breakdown_ms": {k: round(v["ms"], 3) for k, v in prof.items()},
                    "breakdown_tflops": {k: (round(v["flops"] / (v["ms"] * 1e-3) / 1e12, 3) if v["ms"] > 0 else 0.0) for k, v in prof.items()}}
        parity = bench_batch_parity(eng, pred_d, batch, cfg) if (world == 1 and args.batch == 32 and args.seed == 0) else None
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            ref = CpuReference(sd)
            r = ref.run(CPU_SAMPLES["baseline"], args.seed)
            cpu = {"value": r["distinct"] / r["seconds"], "value_slots": r["slots"] / r["seconds"], "unit": UNIT, "cores": ref.cores,
                   "kind": ref.kind, "sample": ref.describe(CPU_SAMPLES["baseline"], r)}
        print(json.dumps({
            "metric": METRIC, "value": value, "value_slots": value_slots, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "dtype_note": "fp32-class: fp16x2 split operands (x = hi + lo), 3 tcgen05 passes per product, fp32 accumulation; fp32 LayerNorm / softmax / residuals",
            "data": "synthetic",
            "config": {"workload": WORKLOAD, "edge_count": COUNT_NOTE, "wireframes_per_step_per_gpu": N, "sequences_per_step_per_gpu": info["B"],
                       "sequences_decoded_per_gpu": info["B_eff"], "decode_steps": S, "memory_rows": info["R"],
                       "weights": f"synthetic seed {args.seed} recipe diverse (32.3 M params, fp32)",
                       "l2": "no explicit flush: per-step activations + cross-K/V cache are GBs, far larger than the 126 MB L2",
                       "parallelism": f"dp{world} (whole batches per rank; NCCL weight broadcast + all-gather of predictions)"},
            "e2e": {"value": e2e_value, "value_slots": e2e_value * slots_per_step / edges_per_step, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                    "ms_per_step": ms_e / e2e_steps, "steps": e2e_steps},
            "gpu_launches": int(launches), "clocks": clocks, "roofline": roofline, "cpu_baseline": cpu, "parity": parity}))
    eng.close()
    if world > 1:
        dist.destroy_process_group()


# =====================================================================================================================
# The other BASELINE.json configs (faceformer_b200/workloads.py).  Same contract: `value` device-resident and CUDA-event
# timed, `e2e` through host buffers, `roofline` for the regime the workload is in.
# =====================================================================================================================
def _dist_setup():
    import torch
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a GPU: faceformer_b200 has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    return torch, dist, world, rank, local, dev


def _timed(torch, dist, world, dev, fn, k):
    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    out = None
    for _ in range(k):
        out = fn()
    e1.record()
    barrier()
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    return float(ms.item()), out


def run_small(args, name):
    """ours_n1 / seq2seq_n1_64: one wireframe per model(batch) call -- the launch- and HBM-bound regime."""
    torch, dist, world, rank, local, dev = _dist_setup()
    from faceformer_b200 import workloads as wl
    from faceformer_b200.engine import Engine
    from faceformer_b200.lib import FFB_OPT_TIMING
    w = wl.make_small(name, args, torch, dev)
    cfg, mode = w["cfg"], w["mode"]
    eng = Engine(cfg, mode, local)
    eng.load_state_dict(w["sd"])
    if args.pdl >= 0:
        from faceformer_b200.lib import FFB_OPT_PDL
        eng.set_option(FFB_OPT_PDL, args.pdl)
    if args.persistent >= 0:
        from faceformer_b200.lib import FFB_OPT_PERSISTENT
        eng.set_option(FFB_OPT_PERSISTENT, args.persistent)
    T = cfg.seq_len(mode)

    def step_device():
        tot = 0
        for c, m, ni in w["calls"]:
            _, s = eng.forward_eval(c, m, ni)
            tot += s
        return tot

    def step_host():
        tot = 0
        for c, m, ni in w["host_calls"]:
            _, s = eng.forward_eval(c, m, ni)
            tot += s
        return tot

    for _ in range(max(1, args.warmup)):
        step_device()
    # per-call geometry and the decode-only device time (FFB_OPT_TIMING) for the HBM roofline of the decode step
    eng.set_option(FFB_OPT_TIMING, 1)
    edges = slots = 0
    bytes_total = dec_ms = enc_ms = 0.0
    steps_total = 0
    for (c, m, ni), b in zip(w["calls"], w["batches"]):
        _, s = eng.forward_eval(c, m, ni)
        torch.cuda.synchronize()
        info = eng.batch_info()
        e_ms, d_ms = eng.phase_times()
        enc_ms += e_ms
        dec_ms += d_ms
        edges += info["B_eff"] * s
        slots += info["B"] * s
        steps_total += s
        bytes_total += s * wl.decoder_step_bytes(cfg, info["R"], info["B_eff"])
    eng.set_option(FFB_OPT_TIMING, 0)
    persistent = bool(eng.used_persistent())                 # (of the last call; seq2seq_n1_64 has one call per step)
    sampler = ClockSampler(local)
    sampler.start()
    l0 = eng.kernel_launches()
    ms, _ = _timed(torch, dist, world, dev, step_device, args.steps)
    launches = eng.kernel_launches() - l0
    clocks = sampler.stop()
    step_host()
    e2e_steps = max(1, min(args.steps, 3))
    ms_e, _ = _timed(torch, dist, world, dev, step_host, e2e_steps)
    from faceformer_b200.lib import FFB_OPT_PROFILE
    eng.set_option(FFB_OPT_PROFILE, 1)                       # event pair around every launch: kernel time without the launch gaps
    step_device()
    prof = eng.profile_read()
    eng.set_option(FFB_OPT_PROFILE, 0)
    peaks = load_peaks()
    value = edges * args.steps / (ms / 1e3)
    h2d = sum(c.nbytes + m.nbytes + (0 if ni is None else ni.nbytes) for c, m, ni in w["host_calls"])
    d2h = sum((int(b["num_input"].max()) if mode == 0 else 1) * T * 8 for b in w["batches"])
    ach = bytes_total / (dec_ms * 1e-3) / 1e9
    line = {
        "metric": METRIC, "value": value, "value_slots": slots * args.steps / (ms / 1e3), "unit": UNIT, "n_gpus": 1, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        "config": {"workload": ("configs/seq2seq.yml greedy decode, 1 wireframe x 64 edges, T=259 (BASELINE configs[0])" if name == "seq2seq_n1_64" else
                                "configs/ours.yml greedy decode, batch=1 (the reference's test loop, trainer.py:51), 16 wireframes n_edges~U[24,216] per step"),
                   "model_calls_per_step": len(w["calls"]), "decode_steps_per_step": steps_total, "edge_count": COUNT_NOTE,
                   "ms_per_decode_step": dec_ms / steps_total, "encode_ms_per_step": enc_ms,
                   "l2": "weights (77 MB as fp16x2 operand pairs) stay resident in the 126 MB L2 between decode steps: DRAM traffic is below the algorithmic bytes"},
        "e2e": {"value": edges * e2e_steps / (ms_e / 1e3), "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                "ms_per_step": ms_e / e2e_steps, "steps": e2e_steps},
        "gpu_launches": int(launches), "clocks": clocks,
        "roofline": {"bound": "hbm", "kernel": ("pd::decode_persistent_kernel (the whole greedy loop in ONE cooperative launch; bound by ~61 grid-wide "
                                                "barriers and dependent L2 round trips per decode step, not by bytes)" if persistent else
                                                "decode step (all kernels of one greedy step; launch-latency bound at this size)"),
                     "persistent_kernel": persistent,
                     "achieved": ach, "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": ach / peaks["hbm_gbs"],
                     "traffic": (load_traffic("decode_persistent_kernel")[0] if persistent and name == "seq2seq_n1_64" else None),
                     "traffic_note": (load_traffic("decode_persistent_kernel")[1] if persistent and name == "seq2seq_n1_64" else None),
                     "algorithmic_bytes_per_decode_step": bytes_total / steps_total, "launches_per_decode_step": launches / args.steps / steps_total,
                     "kernel_ms_by_class": {k: round(v["ms"], 3) for k, v in prof.items()},
                     "kernel_ms_sum_vs_step_ms": [round(sum(v["ms"] for v in prof.values()), 3), round(ms / args.steps, 3)],
                     "peak_source": f"MEASURED_PEAKS.json hbm_gbs ({peaks['source']})"},
        "cpu_baseline": None}
    print(json.dumps(line))
    eng.close()


def run_beam4(args):
    """BASELINE configs[3]: ours-perspective.yml geometry, batch 64, beam width 4.  PARITY UNPINNED (no reference beam search)."""
    torch, dist, world, rank, local, dev = _dist_setup()
    from faceformer_b200.config import OURS_PERSPECTIVE
    from faceformer_b200.engine import Engine
    from faceformer_b200.lib import FFB_OPT_BEAM, FFB_OPT_PROFILE
    cfg, mode, N, W = OURS_PERSPECTIVE, MODE_PARALLEL, args.batch if args.batch != 32 else 64, args.beam
    eng = Engine(cfg, mode, local)
    eng.load_state_dict(synth.synth_state_dict(cfg, mode, args.seed, "diverse"))
    eng.set_option(FFB_OPT_BEAM, W)
    batch = synth.synth_batch(cfg, mode, N, seed=args.seed)
    T = cfg.max_face_length
    coords_h = torch.from_numpy(batch["input"].reshape(N, cfg.num_lines, -1)).pin_memory()
    mask_h = torch.from_numpy(batch["input_mask"].astype(np.uint8)).pin_memory()
    ni_h = torch.from_numpy(batch["num_input"]).pin_memory()
    coords_d, mask_d, ni_d = coords_h.to(dev), mask_h.to(dev), ni_h.to(dev)
    F = int(batch["num_input"].max())
    pred_d = torch.empty((N, F, T), dtype=torch.int64, device=dev)
    pred_h = torch.empty((N, F, T), dtype=torch.int64).pin_memory()

    def step_device():
        return eng.forward_eval(coords_d, mask_d, ni_d, out=pred_d)[1]

    def step_host():
        return eng.forward_eval(coords_h.numpy(), mask_h.numpy(), ni_h.numpy(), out=pred_h.numpy())[1]

    for _ in range(max(1, args.warmup)):
        S = step_device()
    info = eng.batch_info()
    sampler = ClockSampler(local)
    sampler.start()
    l0 = eng.kernel_launches()
    ms, S = _timed(torch, dist, world, dev, step_device, args.steps)
    launches = eng.kernel_launches() - l0
    clocks = sampler.stop()
    step_host()
    e2e_steps = max(1, min(args.steps, 3))
    ms_e, _ = _timed(torch, dist, world, dev, step_host, e2e_steps)
    eng.set_option(FFB_OPT_PROFILE, 1)
    step_device()
    prof = eng.profile_read()
    eng.set_option(FFB_OPT_PROFILE, 0)
    peaks = load_peaks()
    lin = prof["linear_tc"]
    ach = lin["flops"] / (lin["ms"] * 1e-3) / 1e12 if lin["ms"] > 0 else 0.0
    anchors = info["B_eff"] // W                      # distinct anchor sequences; every one carries W hypotheses through every step
    value = anchors * S * args.steps / (ms / 1e3)
    print(json.dumps({
        "metric": METRIC, "value": value, "value_hypotheses": value * W, "value_slots": info["B"] * S * args.steps / (ms / 1e3), "unit": UNIT,
        "n_gpus": 1, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"configs/ours-perspective.yml geometry (num_lines=202, T=38), batch={N}, beam={W} (BASELINE configs[3])",
                   "parity": "UNPINNED: the reference has no beam search; semantics specified in oracle/beam_oracle.py (beam=1 == the reference's greedy loop, tested)",
                   "edge_count": "value counts distinct ANCHOR sequences x steps (each anchor decodes `beam` hypotheses per step: value_hypotheses)",
                   "anchors_decoded": anchors, "hypotheses_decoded": info["B_eff"], "decode_steps": S, "memory_rows": info["R"]},
        "e2e": {"value": anchors * S * e2e_steps / (ms_e / 1e3), "unit": UNIT, "h2d_bytes_per_step": int(coords_h.numel() * 4 + mask_h.numel() + ni_h.numel() * 8),
                "d2h_bytes_per_step": int(pred_h.numel() * 8), "ms_per_step": ms_e / e2e_steps, "steps": e2e_steps},
        "gpu_launches": int(launches), "clocks": clocks,
        "roofline": {"bound": "tensor", "kernel": "tc::gemm_kernel<2> (tcgen05 fp16x2 split, 3 MMA passes)", "achieved": ach, "peak": peaks["bf16_sustained"] / 3,
                     "unit": "TFLOP/s", "frac": ach / (peaks["bf16_sustained"] / 3), "traffic": None,
                     "breakdown_ms": {k: round(v["ms"], 3) for k, v in prof.items()}},
        "cpu_baseline": None}))
    eng.close()


def run_encoder2048(args):
    """BASELINE configs[4]: encoder only (embedding + 6 encoder layers + final norm + the cross K / V cache), 2048-edge wireframes."""
    torch, dist, world, rank, local, dev = _dist_setup()
    from faceformer_b200 import workloads as wl
    from faceformer_b200.engine import Engine
    from faceformer_b200.lib import FFB_OPT_ENCODE_ONLY, FFB_OPT_ENCODER_PRECISION, FFB_OPT_PROFILE
    cfg, mode = OURS.replace(num_lines=2048), MODE_PARALLEL
    N = args.batch if args.batch != 32 else 256
    eng = Engine(cfg, mode, local)
    eng.load_state_dict(synth.synth_state_dict(cfg, mode, 6, "diverse"))
    eng.set_option(FFB_OPT_ENCODER_PRECISION, 0)        # throughput mode: fp16x2 tcgen05 GEMMs + attention (the float64 encoder is for <= 320-row wireframes)
    eng.set_option(FFB_OPT_ENCODE_ONLY, 1)
    one = synth.synth_batch(cfg, mode, 4, seed=8, num_edges=np.full(4, 2048, np.int64))
    reps = (N + 3) // 4
    coords_h = torch.from_numpy(np.tile(one["input"].reshape(4, cfg.num_lines, -1), (reps, 1, 1))[:N]).pin_memory()
    mask_h = torch.from_numpy(np.tile(one["input_mask"].astype(np.uint8), (reps, 1))[:N]).pin_memory()
    ni_h = torch.from_numpy(np.tile(one["num_input"], reps)[:N]).pin_memory()
    coords_d, mask_d, ni_d = coords_h.to(dev), mask_h.to(dev), ni_h.to(dev)

    def step_device():
        eng.encode(coords_d, mask_d, ni_d)

    def step_host():
        eng.encode(coords_h.numpy(), mask_h.numpy(), ni_h.numpy())
        torch.cuda.synchronize()

    for _ in range(max(1, args.warmup)):
        step_device()
    info = eng.batch_info()
    sampler = ClockSampler(local)
    sampler.start()
    l0 = eng.kernel_launches()
    ms, _ = _timed(torch, dist, world, dev, step_device, args.steps)
    launches = eng.kernel_launches() - l0
    clocks = sampler.stop()
    e2e_steps = max(1, min(args.steps, 3))
    ms_e, _ = _timed(torch, dist, world, dev, step_host, e2e_steps)
    eng.set_option(FFB_OPT_PROFILE, 1)
    step_device()
    prof = eng.profile_read()
    eng.set_option(FFB_OPT_PROFILE, 0)
    peaks = load_peaks()
    flops = wl.encoder_flops(cfg, [cfg.mem_len] * N)
    ach = flops / (ms / args.steps * 1e-3) / 1e12
    peak = peaks["bf16_sustained"] / 3
    print(json.dumps({
        "metric": "encoder_wireframes_per_sec", "value": N * args.steps / (ms / 1e3), "unit": "wireframes/s", "n_gpus": 1, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "dtype_note": "fp16x2 split operands, 3 tcgen05 passes per product, fp32 accumulation / LayerNorm / softmax", "data": "synthetic",
        "config": {"workload": f"encoder only, synthetic 2048-edge wireframes (L=2052), batch={N} (BASELINE configs[4])", "memory_rows": info["R"],
                   "algorithmic_tflop_per_step": flops / 1e12, "fp16_fallbacks": eng.fp16_fallbacks(),
                   "l2": "activations of one batch are GBs, far larger than the 126 MB L2"},
        "e2e": {"value": N * e2e_steps / (ms_e / 1e3), "unit": "wireframes/s", "h2d_bytes_per_step": int(coords_h.numel() * 4 + mask_h.numel() + ni_h.numel() * 8),
                "d2h_bytes_per_step": 0, "ms_per_step": ms_e / e2e_steps, "steps": e2e_steps,
                "note": "the encoder's product (memory + K/V cache) stays on the device for the decode; nothing is read back"},
        "gpu_launches": int(launches), "clocks": clocks,
        "roofline": {"bound": "tensor", "kernel": "whole encoder step (tc::gemm_kernel<2> + attention)", "achieved": ach, "peak": peak, "unit": "TFLOP/s",
                     "frac": ach / peak, "traffic": None, "mma_passes": 3,
                     "breakdown_ms": {k: round(v["ms"], 3) for k, v in prof.items()},
                     "breakdown_tflops": {k: (round(v["flops"] / (v["ms"] * 1e-3) / 1e12, 3) if v["ms"] > 0 else 0.0) for k, v in prof.items()}},
        "cpu_baseline": None}))
    eng.close()


def run_split507(args):
    """BASELINE configs[2]: the synthetic ours.yml test split (507 wireframes), global batch 128, every global batch split over ALL ranks
    (strong scaling: total work fixed), predictions all-gathered; tensor-identical to the single-GPU batch-128 run."""
    torch, dist, world, rank, local, dev = _dist_setup()
    from faceformer_b200 import sharding
    from faceformer_b200.engine import Engine, pack_state_dict
    cfg, mode = OURS, MODE_PARALLEL
    T = cfg.max_face_length
    eng = Engine(cfg, mode, local)
    if rank == 0:
        blob = torch.from_numpy(pack_state_dict(synth.synth_state_dict(cfg, mode, args.seed, "diverse"), cfg, mode)).to(dev)
    else:
        blob = torch.empty(eng.weight_count(), dtype=torch.float32, device=dev)
    sharding.broadcast_weights(blob, 0)
    eng.load_blob(blob)
    n_total, gb = args.split_size, 128
    ne = synth.synth_num_edges(cfg, n_total, args.seed + 1)
    batches = []
    for s0 in range(0, n_total, gb):
        b = synth.synth_batch(cfg, mode, len(ne[s0:s0 + gb]), seed=args.seed + 1 + s0, num_edges=ne[s0:s0 + gb])
        n = len(b["num_input"])
        batches.append((torch.from_numpy(b["input"].reshape(n, cfg.num_lines, -1)).to(dev), torch.from_numpy(b["input_mask"].astype(np.uint8)).to(dev),
                        torch.from_numpy(b["num_input"]).to(dev)))
    split = sharding.SplitDecoder(eng)
    split.connect()
    steps_seen = []

    def step_device():
        steps_seen.clear()
        for c, m, ni in batches:
            _, s = split.forward_eval(c, m, ni, gather=True)
            steps_seen.append(s)
        return list(steps_seen)

    for _ in range(max(1, args.warmup)):
        S = step_device()
    # edges of the whole split: distinct sequences (+1 padded-anchor copy per wireframe shorter than its batch's F is NOT counted) x steps
    edges = slots = 0
    for (c, m, ni), s in zip(batches, S):
        n = ni.cpu().numpy()
        edges += int(n.sum()) * s
        slots += len(n) * int(n.max()) * s
    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start()
    l0 = eng.kernel_launches()
    ms, S = _timed(torch, dist, world, dev, step_device, args.steps)
    launches = eng.kernel_launches() - l0
    clocks = sampler.stop() if sampler else None
    parts = [sharding.split_batch(ni.cpu().numpy(), world) for _, _, ni in batches]
    load = np.zeros(world)
    for (c, m, ni), p in zip(batches, parts):
        n = ni.cpu().numpy()
        for r in range(world):
            load[r] += (n[p[r]] + 1).sum()
    if rank == 0:
        value = edges * args.steps / (ms / 1e3)
        print(json.dumps({
            "metric": METRIC, "value": value, "value_slots": slots * args.steps / (ms / 1e3), "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic",
            "config": {"workload": f"configs/ours.yml synthetic test split: {n_total} wireframes, n_edges~U[24,216], global batch {gb}, every global batch split over "
                                   f"{world} GPU(s) (BASELINE configs[2])", "edge_count": COUNT_NOTE, "global_batches": len(batches), "decode_steps": S,
                       "parallelism": f"dp{world}: wireframes of one batch dealt LPT to ranks; F fixed per global batch; stop predicate exchanged per step through "
                                      "peer-mapped flag words (CUDA IPC over NVLink); NCCL all-gather of predictions per global batch",
                       "load_imbalance_max_over_mean": float(load.max() / load.mean())},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0,
                    "note": "device-resident split (the all-gather is inside the timed region); the host-buffer e2e of one batch is the default workload's"},
            "gpu_launches": int(launches), "clocks": clocks,
            "roofline": {"bound": "tensor", "kernel": "see the default workload; this line measures strong scaling", "achieved": None, "peak": None, "unit": "TFLOP/s",
                         "frac": None, "traffic": None},
            "cpu_baseline": None}))
    eng.close()
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=32)
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--tc-format", type=int, default=2, choices=[2, 3], help="tensor-core operand format: 2 fp16x2, 3 bf16x3")
    ap.add_argument("--tc", type=int, default=1, choices=[0, 1, 2],
                    help="decode-step linear layers: 0 fp32 SIMT, 1 auto (tcgen05 split-precision GEMM at every M), 2 force tcgen05")
    ap.add_argument("--pdl", type=int, default=-1, help="override FFB_OPT_PDL (-1 = library default)")
    ap.add_argument("--gemm-variant", type=int, default=-1, help="override FFB_OPT_GEMM_VARIANT (-1 = library default)")
    ap.add_argument("--attn-x", type=int, default=-1, help="override FFB_OPT_ATTN_X (bit mask; -1 = library default)")
    ap.add_argument("--workload", default="ours32", choices=["ours32", "ours_n1", "seq2seq_n1_64", "split507", "beam4", "encoder2048"],
                    help="ours32 = BASELINE configs[1] (the headline, default); the others are the remaining BASELINE configs (faceformer_b200/workloads.py)")
    ap.add_argument("--persistent", type=int, default=-1, help="override FFB_OPT_PERSISTENT (0 off, 1 auto, 2 wherever supported; -1 = library default)")
    ap.add_argument("--beam", type=int, default=4, help="beam width of --workload beam4")
    ap.add_argument("--split-size", type=int, default=507, help="wireframes of --workload split507")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    elif args.workload in ("ours_n1", "seq2seq_n1_64"):
        run_small(args, args.workload)
    elif args.workload == "beam4":
        run_beam4(args)
    elif args.workload == "encoder2048":
        run_encoder2048(args)
    elif args.workload == "split507":
        run_split507(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
