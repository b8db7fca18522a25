"""The persistent cooperative decode kernel (csrc/persist.cuh, FFB_OPT_PERSISTENT): the whole greedy loop in one launch.
Same bar as every other pipeline: tokens exact against the reference goldens, executed steps equal (the stop predicate is evaluated on the
device), last-step logits within the logit tolerance, seq2seq extras; and agreement with the multi-kernel path on fresh batches."""
import numpy as np
import pytest
import torch

from faceformer_b200 import synth
from faceformer_b200.config import MODE_PARALLEL, MODE_SEQ2SEQ, OURS, SEQ2SEQ
from faceformer_b200.engine import Engine
from faceformer_b200.lib import FFB_OPT_PERSISTENT
from oracle import faceformer_oracle as orc
from util import LOGIT_TOL, load_case, logits_close

pytestmark = pytest.mark.gpu


def decode(g, batch, persist):
    e = Engine(g["cfg"], g["mode"], 0)
    e.load_state_dict(g["sd"])
    e.set_option(FFB_OPT_PERSISTENT, persist)
    coords = torch.from_numpy(batch["input"]).cuda().flatten(2)
    mask, ni = torch.from_numpy(batch["input_mask"]).cuda(), torch.from_numpy(batch["num_input"]).cuda()
    pred, steps = e.forward_eval(coords, mask, ni)
    out = dict(pred=pred.cpu().numpy(), steps=steps, logits=e.get_last_logits().cpu().numpy(), used=e.used_persistent(), launches=e.kernel_launches())
    if g["mode"] == MODE_SEQ2SEQ:
        out["pointer"] = e.get_last_pointer().cpu().numpy()
    e.close()
    return out


def test_seq2seq_single_wireframe_takes_the_persistent_kernel_by_default():
    """BASELINE configs[0]: one 64-edge wireframe, 258 steps -- auto mode must choose the persistent kernel and match the reference."""
    g = load_case("seq2seq_single64")
    r = decode(g, g["batch"], 1)
    assert r["used"]
    assert r["steps"] == g["steps"]
    assert np.array_equal(r["pred"], g["predict"]), f"{(r['pred'] != g['predict']).sum()} token mismatches"
    ok, d = logits_close(r["logits"], g["last_logits"], b64=g.get("last_logits64"))
    assert ok, f"last-step logits differ by {d}"
    off = decode(g, g["batch"], 0)
    assert not off["used"] and np.array_equal(off["pred"], g["predict"])
    assert r["pointer"].shape == off["pointer"].shape
    assert np.max(np.abs(r["pointer"] - off["pointer"])) <= LOGIT_TOL        # 'pointer' extra (model.py:216-217)
    assert r["launches"] < off["launches"] / 50                               # one launch instead of ~72 per step


@pytest.mark.parametrize("name", ["ours_parallel_small", "perspective_small", "ours_wide300", "mid_parallel_trained", "mid_parallel_trained_b"])
def test_forced_persistent_kernel_matches_the_reference_goldens(name):
    g = load_case(name)
    r = decode(g, g["batch"], 2)
    assert r["used"]
    assert r["steps"] == g["steps"]
    assert np.array_equal(r["pred"], g["predict"]), f"{(r['pred'] != g['predict']).sum()} token mismatches"
    ok, d = logits_close(r["logits"], g["last_logits"], b64=g.get("last_logits64"))
    assert ok, f"last-step logits differ by {d}"


@pytest.mark.parametrize("mode,n,seed,lo,hi", [(MODE_PARALLEL, 1, 31, 24, 24), (MODE_PARALLEL, 1, 32, 7, 7), (MODE_PARALLEL, 5, 33, 3, 12),
                                               (MODE_SEQ2SEQ, 3, 34, 5, 40)])
def test_persistent_and_multi_kernel_paths_agree_on_fresh_batches(mode, n, seed, lo, hi):
    """One wireframe per batch (the reference's test loop, trainer.py:51), a ragged small batch and a seq2seq batch: both pipelines are
    fp32-class evaluations of the same function -- same tokens, same steps, logits within the tolerance."""
    cfg = OURS if mode == MODE_PARALLEL else load_case("seq2seq_single64")["cfg"]
    g = dict(cfg=cfg, mode=mode, sd=synth.synth_state_dict(cfg, mode, seed, "diverse"))
    batch = synth.synth_batch(cfg, mode, n, seed, lo=lo, hi=hi)
    a, b = decode(g, batch, 2), decode(g, batch, 0)
    assert a["used"] and not b["used"]
    assert a["steps"] == b["steps"]
    assert np.array_equal(a["pred"], b["pred"]), f"{(a['pred'] != b['pred']).sum()} token mismatches"
    ok, d = logits_close(a["logits"], b["logits"])
    assert ok, d


def test_auto_mode_takes_the_persistent_kernel_only_below_the_measured_crossover():
    """FFB_OPT_PERSISTENT = 1 (default): sequences x (T - 1) <= 896 decoder rows (profiles/probe_persist_threshold_r2.json)."""
    cfg = OURS
    g = dict(cfg=cfg, mode=MODE_PARALLEL, sd=synth.synth_state_dict(cfg, MODE_PARALLEL, 5, "diverse"))
    small = decode(g, synth.synth_batch(cfg, MODE_PARALLEL, 1, 35, lo=7, hi=7), 1)        # 7 x 36 = 252 rows
    large = decode(g, synth.synth_batch(cfg, MODE_PARALLEL, 1, 36, lo=40, hi=40), 1)      # 40 x 36 = 1440 rows
    assert small["used"] and not large["used"]
    assert small["launches"] < large["launches"] / 10


def test_hybrid_mode_hands_over_to_the_per_step_kernels_at_the_measured_crossover():
    """FFB_OPT_PERSISTENT = 3: the persistent kernel runs the steps whose row count (sequences x prefix length) is <= 896
    (profiles/probe_persist_threshold_r2.json) and the per-step kernels continue from the same token / state buffers: a 7-edge wireframe is
    decoded entirely inside it (7 x 36 = 252 rows), a 40-edge one for its first 22 steps, a 120-edge one for 7 -- all token-identical to the
    multi-kernel decode."""
    cfg = OURS
    g = dict(cfg=cfg, mode=MODE_PARALLEL, sd=synth.synth_state_dict(cfg, MODE_PARALLEL, 5, "diverse"))
    prev = None
    for seed, n in ((35, 7), (36, 40), (37, 120)):
        batch = synth.synth_batch(cfg, MODE_PARALLEL, 1, seed, lo=n, hi=n)
        a, b = decode(g, batch, 3), decode(g, batch, 0)
        assert a["used"] and not b["used"]
        assert a["steps"] == b["steps"] and np.array_equal(a["pred"], b["pred"]), f"{n} edges: {(a['pred'] != b['pred']).sum()} token mismatches"
        ok, d = logits_close(a["logits"], b["logits"])
        assert ok, d
        assert a["launches"] < b["launches"]
        if prev is not None:
            assert a["launches"] > prev                     # fewer steps inside the persistent kernel as the wireframe grows
        prev = a["launches"]


@pytest.mark.parametrize("cfg,mode,n,seed,lo,hi,steps", [(SEQ2SEQ, MODE_SEQ2SEQ, 2, 68, 6, 14, 4),      # cumulative EOS count reaches N at step 4 (model.py:207-210)
                                                         (SEQ2SEQ, MODE_SEQ2SEQ, 2, 62, 6, 14, 2),
                                                         (OURS, MODE_PARALLEL, 1, 108, 5, 9, 5),        # all sequences emit a special token at step 5 (model_para.py:232)
                                                         (OURS, MODE_PARALLEL, 1, 100, 5, 9, 1)])
def test_early_stop_is_decided_on_the_device_like_the_reference_does(cfg, mode, n, seed, lo, hi, steps):
    """The persistent kernel evaluates the stop predicate itself after every step and leaves the loop: executed steps, tokens up to the stop
    and the zero padding behind it must equal the numpy oracle's (= the reference's loop), on full-size (E = 512) models."""
    sd = synth.synth_state_dict(cfg, mode, seed, "diverse")
    batch = synth.synth_batch(cfg, mode, n, seed, lo=lo, hi=hi)
    want = orc.forward_eval(sd, cfg.to_dict(), mode, batch, return_trace=True, max_steps=steps + 3)
    assert want["steps"] == steps
    g = dict(cfg=cfg, mode=mode, sd=sd)
    a, b = decode(g, batch, 1), decode(g, batch, 0)
    assert a["used"] and not b["used"]
    for r in (a, b):
        assert r["steps"] == steps
        assert np.array_equal(r["pred"], want["predict"]), f"{(r['pred'] != want['predict']).sum()} token mismatches"
    ok, d = logits_close(a["logits"], want["logits"][-1])
    assert ok, d


def test_hybrid_mode_early_stop_inside_the_persistent_part():
    """30 edges: the persistent kernel may run 29 of the 36 steps (mode 3); the batch stops at step 2 inside it, so the 7 steps that were
    launched behind it on the per-step kernels must all exit on the device-side stop flag -- tokens, steps and padding as the oracle's loop."""
    cfg, seed = OURS, 216
    sd = synth.synth_state_dict(cfg, MODE_PARALLEL, seed, "diverse")
    batch = synth.synth_batch(cfg, MODE_PARALLEL, 1, seed, lo=30, hi=30)
    want = orc.forward_eval(sd, cfg.to_dict(), MODE_PARALLEL, batch, return_trace=True, max_steps=5)
    assert want["steps"] == 2
    r = decode(dict(cfg=cfg, mode=MODE_PARALLEL, sd=sd), batch, 3)
    assert r["used"] and r["steps"] == 2
    assert np.array_equal(r["pred"], want["predict"])
    ok, d = logits_close(r["logits"], want["logits"][-1])
    assert ok, d
