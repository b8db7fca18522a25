"""Parity of the CUDA path (through the C ABI) with the reference's outputs (golden fixtures) and with the
numpy oracle on fresh seeded inputs.  Tokens: exact.  Pointer logits / memory: within 1e-4 (north_star)."""
import numpy as np
import pytest
import torch

from faceformer_b200 import synth
from faceformer_b200.config import MODE_PARALLEL, MODE_SEQ2SEQ, OURS, TINY
from faceformer_b200.engine import Engine
from faceformer_b200.lib import FFB_OPT_DEDUP_PAD, FFB_OPT_PRUNE_LAST, FFBError
from oracle import faceformer_oracle as orc
from util import CASES_ALL, LOGIT_TOL, load_case, logits_close, valid_rows_mask

pytestmark = pytest.mark.gpu


def make_engine(g):
    e = Engine(g["cfg"], g["mode"], 0)
    e.load_state_dict(g["sd"])
    return e


def run(e, batch, device_path):
    b = batch
    if device_path:
        coords = torch.from_numpy(b["input"]).cuda().flatten(2)
        mask = torch.from_numpy(b["input_mask"]).cuda()
        ni = torch.from_numpy(b["num_input"]).cuda()
    else:
        coords, mask, ni = b["input"].reshape(b["input"].shape[0], b["input"].shape[1], -1), b["input_mask"], b["num_input"]
    pred, steps = e.forward_eval(coords, mask, ni)
    if device_path:
        pred = pred.cpu().numpy()
    return pred, steps


@pytest.mark.parametrize("name", CASES_ALL)
@pytest.mark.parametrize("device_path", [True, False])
def test_golden_full_decode(name, device_path):
    g = load_case(name)
    e = make_engine(g)
    pred, steps = run(e, g["batch"], device_path)
    assert pred.dtype == np.int64 and pred.shape == g["predict"].shape
    assert steps == g["steps"]
    assert np.array_equal(pred, g["predict"]), f"{(pred != g['predict']).sum()} token mismatches"
    lg = e.get_last_logits()
    lg = lg.cpu().numpy() if device_path else lg
    ok, d = logits_close(lg, g["last_logits"], b64=g.get("last_logits64"))
    assert ok, f"last-step logits differ by {d}"
    mem = e.get_memory()
    mem = mem.cpu().numpy() if device_path else mem
    vm = valid_rows_mask(g["batch"], g["cfg"])
    assert np.max(np.abs(mem[vm] - g["memory"][vm])) <= LOGIT_TOL
    assert np.all(mem[~vm] == 0)
    assert e.kernel_launches() > 0
    e.close()


def test_golden_encoder_2048_edges():
    """BASELINE.json configs[4] geometry: 2048-edge wireframe (L = 2052 memory rows, flash-style key streaming in the encoder
    self-attention); encoder memory against the reference's (every 8th row kept in the fixture)."""
    g = load_case("encoder_2048")
    e = make_engine(g)
    b = g["batch"]
    e.encode(torch.from_numpy(b["input"]).cuda().flatten(2), torch.from_numpy(b["input_mask"]).cuda(), torch.from_numpy(b["num_input"]).cuda())
    mem = e.get_memory().cpu().numpy()
    assert mem.shape == (1, g["cfg"].mem_len, g["cfg"].num_model)
    assert np.max(np.abs(mem[:, g["rows"]] - g["memory_rows"])) <= LOGIT_TOL
    e.close()


@pytest.mark.parametrize("name", CASES_ALL)
def test_golden_forced_prefix_logits(name):
    g = load_case(name)
    e = make_engine(g)
    e.set_option(FFB_OPT_DEDUP_PAD, 0)
    b = g["batch"]
    e.encode(b["input"].reshape(b["input"].shape[0], b["input"].shape[1], -1), b["input_mask"], b["num_input"])
    lg = e.forced_prefix_logits(g["prefix"])
    ok, d = logits_close(lg, g["prefix_logits"], b64=g.get("prefix_logits64"))
    assert ok, f"forced-prefix logits differ by {d}"
    e.close()


@pytest.mark.parametrize("name", ["tiny_parallel_trained_b", "tiny_parallel_ragged", "ours_parallel_small"])
def test_dedup_and_pruning_do_not_change_results(name):
    g = load_case(name)
    res = []
    for dedup, prune in [(1, 1), (0, 1), (1, 0), (0, 0)]:
        e = make_engine(g)
        e.set_option(FFB_OPT_DEDUP_PAD, dedup)
        e.set_option(FFB_OPT_PRUNE_LAST, prune)
        pred, steps = run(e, g["batch"], True)
        res.append((pred, steps, e.get_last_logits().cpu().numpy(), e.batch_info()))
        e.close()
    for pred, steps, lg, info in res:
        assert steps == g["steps"] and np.array_equal(pred, g["predict"])
        assert logits_close(lg, res[0][2], 5e-5)[0]        # pruned / un-pruned last layer take different kernels (attn_last vs tcgen05 attention)
    assert res[1][3]["B_eff"] == res[1][3]["B"] and res[0][3]["B_eff"] <= res[0][3]["B"]


@pytest.mark.parametrize("mode,n,seed,lo,hi", [(MODE_PARALLEL, 6, 21, 1, 28), (MODE_PARALLEL, 3, 22, 20, 28),
                                               (MODE_SEQ2SEQ, 4, 23, 2, 28), (MODE_SEQ2SEQ, 1, 24, 9, 9)])
def test_fresh_seeds_against_oracle(mode, n, seed, lo, hi):
    cfg = TINY
    sd = synth.synth_state_dict(cfg, mode, seed, "diverse")
    batch = synth.synth_batch(cfg, mode, n, seed, lo=lo, hi=hi)
    want = orc.forward_eval(sd, cfg.to_dict(), mode, batch, return_trace=True)
    e = Engine(cfg, mode, 0)
    e.load_state_dict(sd)
    pred, steps = run(e, batch, True)
    assert steps == want["steps"]
    assert np.array_equal(pred, want["predict"])
    ok, d = logits_close(e.get_last_logits().cpu().numpy(), want["logits"][-1])
    assert ok, d
    if mode == MODE_SEQ2SEQ:                                    # model.py:216-217 extras
        ptr = e.get_last_pointer().cpu().numpy()
        assert ptr.shape == want["pointer"].shape
        assert np.max(np.abs(ptr - want["pointer"])) <= LOGIT_TOL
    e.close()


def test_trained_fixture_against_oracle_many_wireframes():
    """Non-degenerate weights: 24 polygon wireframes, token-exact (incl. post-EOS tokens)."""
    g = load_case("tiny_parallel_trained")
    batch = synth.polygon_batch(g["cfg"], 24, seed=99)
    want = orc.forward_eval(g["sd"], g["cfg"].to_dict(), g["mode"], batch, return_trace=True)
    e = make_engine(g)
    pred, steps = run(e, batch, True)
    assert steps == want["steps"]
    assert np.array_equal(pred, want["predict"])
    assert len(np.unique(pred)) > 10
    e.close()


def test_unsupported_inputs_fail_loudly():
    g = load_case("tiny_parallel_ragged")
    e = make_engine(g)
    b = {k: v.copy() for k, v in g["batch"].items()}
    coords = b["input"].reshape(b["input"].shape[0], b["input"].shape[1], -1)
    bad = b["input_mask"].copy(); bad[0, 0] = True; bad[0, 5] = False           # hole in the mask
    with pytest.raises(FFBError, match="prefix form"):
        e.encode(coords, bad, b["num_input"])
    ni = b["num_input"].copy(); ni[0] = 28                                      # anchors beyond un-masked rows
    with pytest.raises(FFBError, match="exceeds"):
        e.encode(coords, b["input_mask"], ni)
    with pytest.raises(FFBError):
        Engine(g["cfg"], g["mode"], 0).forward_eval(coords, b["input_mask"], b["num_input"])   # no weights
    e.close()


def test_model_class_boundary():
    """model_class(**cfg.model) + forward(dict) -> dict, strict state_dict load (trainer.py:20,27-28)."""
    from faceformer_b200.models import SurfaceFormer_Parallel_B200
    g = load_case("tiny_parallel_trained")
    cfg = g["cfg"]
    m = SurfaceFormer_Parallel_B200(**cfg.model_kwargs(MODE_PARALLEL), max_num_faces=42, label_seq_length=128).eval()
    m.load_state_dict({k: torch.from_numpy(v) for k, v in g["sd"].items()}, strict=True)
    m = m.cuda()
    batch = {k: torch.from_numpy(v).cuda() for k, v in g["batch"].items()}
    out = m(batch)
    assert out is batch and out["predict"].is_cuda and out["predict"].dtype == torch.int64
    assert np.array_equal(out["predict"].cpu().numpy(), g["predict"])
    with pytest.raises(FFBError, match="no CPU path"):
        m({k: v.cpu() for k, v in batch.items()})
    out = m.train()(batch)                 # teacher-forced forward pass (row f4; tests/test_gpu_train.py holds the parity checks)
    F, T = int(g["batch"]["num_input"].max()), g["cfg"].seq_len(g["mode"])
    assert tuple(out["pointer"].shape) == (len(g["batch"]["num_input"]) * F, T - 1, g["cfg"].num_model)
    assert tuple(out["embedding"].shape)[0] == tuple(out["label"].shape)[0] == tuple(out["pointer"].shape)[0]
    with pytest.raises(NotImplementedError):
        m.forward_train(batch, scheduled_sampling_ratio=0.3)


def test_strict_logit_bar_report():
    """north_star's bar read literally: |logit - reference fp32 logit| <= 1e-4 (x max|logit| / 32 above 32), WITHOUT the noise-floor clause of
    util.logits_close.  Default path, every full-size golden.  The bar holds everywhere except where the reference's own fp32 run is >= 0.8e-4 from
    its float64 evaluation (two correct fp32 evaluations cannot be asked to agree more closely than either agrees with the exact result); those
    fixtures are pinned here BY NAME with their measured distance, so a regression of the decoder stack shows up as a failure, not as drift
    inside a widened tolerance.  Distances to the float64 evaluation are asserted to stay below the reference's own."""
    known_over = {"ours_wide300": 1.5e-4}            # measured 1.26e-4 (fp16x2 pipeline); reference fp32 vs float64 there: 0.86e-4
    report = {}
    for name in ["ours_parallel_small", "seq2seq_single64", "perspective_small", "ours_wide300", "mid_parallel_trained", "mid_parallel_trained_b"]:
        g = load_case(name)
        e = make_engine(g)
        run(e, g["batch"], True)
        lg = e.get_last_logits().cpu().numpy()
        e.close()
        ref, ref64 = g["last_logits"], g["last_logits64"]
        m = ref != np.finfo(np.float32).min
        scale = max(1.0, float(np.max(np.abs(ref[m]))) / 32.0)
        d32 = float(np.max(np.abs(lg[m] - ref[m]))) / scale
        d64 = float(np.max(np.abs(lg[m].astype(np.float64) - ref64[m])))
        n64 = float(np.max(np.abs(ref[m].astype(np.float64) - ref64[m])))
        report[name] = (d32, d64, n64)
        assert d32 <= known_over.get(name, LOGIT_TOL), f"{name}: {d32:.3e} from the reference's fp32 logits (scaled), bar {known_over.get(name, LOGIT_TOL):.1e}"
        assert d64 <= n64, f"{name}: {d64:.3e} from the float64 evaluation, the reference's own fp32 run is {n64:.3e} from it"
    print("strict logit report (scaled |ours - ref32|, |ours - ref64|, |ref32 - ref64|):", {k: tuple(f"{x:.2e}" for x in v) for k, v in report.items()})
