"""The split-precision tensor-core path (tcgen05 bf16x3 GEMM, gemm_tc.cuh) through the C ABI:
op-level accuracy against an fp64 product, and end-to-end parity with the reference's outputs when
EVERY decode step is forced through it."""
import numpy as np
import pytest
import torch

from faceformer_b200.config import MODE_PARALLEL, OURS
from faceformer_b200.engine import Engine
from faceformer_b200.lib import FFB_OPT_DEDUP_PAD, FFB_OPT_TC_FORMAT, FFB_OPT_TENSOR_CORE
from util import LOGIT_TOL, load_case, logits_close, valid_rows_mask

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def eng():
    e = Engine(OURS, MODE_PARALLEL, 0)
    yield e
    e.close()


def _t(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


@pytest.mark.parametrize("M,N,K", [(1, 256, 32), (128, 256, 512), (100, 512, 512), (129, 1536, 512), (1000, 512, 1024),
                                   (4097, 1024, 512), (20000, 512, 512)])
@pytest.mark.parametrize("variant", ["plain", "bias_relu", "bias_res", "via_split"])
@pytest.mark.parametrize("fmt", [2, 3, 21, 23])
def test_tc_linear_is_fp32_class(eng, M, N, K, variant, fmt):
    from faceformer_b200.lib import FFB_OPT_GEMM_VARIANT
    # 21 = fp16x2, pipeline variant 1 (3 stages, 2 staging buffers); 23 = fp16x2 on CTA pairs (gemm_tc2.cuh, cta_group::2)
    eng.set_option(FFB_OPT_GEMM_VARIANT, {21: 1, 23: 3}.get(fmt, 0))
    fmt = 2 if fmt in (21, 23) else fmt
    eng.set_option(FFB_OPT_TC_FORMAT, fmt)
    rng = np.random.default_rng(M + N + K)
    A = (rng.normal(size=(M, K)) * rng.choice([0.01, 1.0, 30.0], size=(M, 1))).astype(np.float32)
    W = (rng.normal(size=(N, K)) / np.sqrt(K)).astype(np.float32)
    b = rng.normal(size=N).astype(np.float32)
    R = rng.normal(size=(M, N)).astype(np.float32)
    ref = A.astype(np.float64) @ W.astype(np.float64).T
    if variant == "plain":
        got = eng.op_linear_tc(_t(A), _t(W))
    elif variant == "bias_relu":
        got = eng.op_linear_tc(_t(A), _t(W), bias=_t(b), relu=True)
        ref = np.maximum(ref + b, 0)
    elif variant == "bias_res":
        got = eng.op_linear_tc(_t(A), _t(W), bias=_t(b), R=_t(R))
        ref = R + (ref + b)
    else:
        got = eng.op_linear_tc(_t(A), _t(W), bias=_t(b), relu=True, via_split=True)
        ref = np.maximum(ref + b, 0)
    got = got.cpu().numpy().astype(np.float64)
    simt = eng.op_linear(_t(A), _t(W)).cpu().numpy().astype(np.float64)
    plain_ref = A.astype(np.float64) @ W.astype(np.float64).T
    # error of each path relative to the row scale of the exact product
    scale = np.abs(plain_ref).max(axis=1, keepdims=True) + 1e-30
    scale_out = np.maximum(scale, np.abs(ref).max(axis=1, keepdims=True))
    if fmt == 2:
        # fp16x2: the low half of a small activation is subnormal in fp16, an ABSOLUTE error floor of ~3e-8 per element
        # (harmless where it matters: these outputs are added to an O(1) residual stream); judge tiny rows against 0.05
        scale_out = np.maximum(scale_out, 0.05)
    err_tc = np.max(np.abs(got - ref) / scale_out)
    err_simt = np.max(np.abs(simt - plain_ref) / scale)
    assert err_tc <= 2e-6, (err_tc, err_simt)                   # fp32-class (one bf16 pass would be ~4e-3)
    assert err_tc <= 4 * err_simt + (2e-7 if fmt == 3 else 1.5e-6), (err_tc, err_simt)  # ~ the fp32 FFMA kernel's own noise


@pytest.mark.parametrize("fmt", [2, 3])
def test_tc_accumulation_has_no_truncation_bias(eng, fmt):
    """All-positive operands: a round-toward-zero accumulator would show a systematic negative error."""
    eng.set_option(FFB_OPT_TC_FORMAT, fmt)
    rng = np.random.default_rng(5)
    M, N, K = 512, 256, 1024
    A = rng.uniform(0.5, 1.5, size=(M, K)).astype(np.float32)
    W = rng.uniform(0.5, 1.5, size=(N, K)).astype(np.float32)
    ref = A.astype(np.float64) @ W.astype(np.float64).T
    got = eng.op_linear_tc(_t(A), _t(W)).cpu().numpy().astype(np.float64)
    rel = (got - ref) / ref
    print("tc accumulation: max |rel| %.3e, mean rel %.3e" % (np.abs(rel).max(), rel.mean()))
    # The tensor core truncates when it adds into its fp32 accumulator: on all-positive data (the worst case) that is a
    # visible negative bias.  Draining to round-to-nearest register accumulators every 2 k-blocks keeps it at ~2e-7
    # (measured -2.1e-7); an undrained K=1024 chain (384 truncating adds) would sit an order of magnitude lower.
    # fp16x2 drains every 4 k-blocks and its two correction products sit between the dominant ones (measured -5.8e-7).
    assert np.abs(rel).max() <= (1e-6 if fmt == 3 else 1.5e-6)
    assert abs(rel.mean()) <= (5e-7 if fmt == 3 else 1e-6), rel.mean()


@pytest.mark.parametrize("name", ["ours_parallel_small", "seq2seq_single64", "perspective_small", "ours_wide300"])
@pytest.mark.parametrize("fmt", [2, 3])
def test_golden_full_decode_forced_tensor_core(name, fmt):
    g = load_case(name)
    e = Engine(g["cfg"], g["mode"], 0)
    e.load_state_dict(g["sd"])
    e.set_option(FFB_OPT_TC_FORMAT, fmt)
    e.set_option(FFB_OPT_TENSOR_CORE, 2)
    b = g["batch"]
    coords = torch.from_numpy(b["input"]).cuda().flatten(2)
    pred, steps = e.forward_eval(coords, torch.from_numpy(b["input_mask"]).cuda(), torch.from_numpy(b["num_input"]).cuda())
    assert steps == g["steps"]
    assert np.array_equal(pred.cpu().numpy(), g["predict"])
    ok, d = logits_close(e.get_last_logits().cpu().numpy(), g["last_logits"], b64=g.get("last_logits64"))
    assert ok, f"last-step logits differ by {d}"
    assert e.kernel_launches() > 0 and e.fp16_fallbacks() == 0
    e.close()


@pytest.mark.parametrize("name", ["ours_parallel_small", "seq2seq_single64"])
@pytest.mark.parametrize("fmt", [2, 3])
def test_golden_forced_prefix_logits_forced_tensor_core(name, fmt):
    g = load_case(name)
    e = Engine(g["cfg"], g["mode"], 0)
    e.load_state_dict(g["sd"])
    e.set_option(FFB_OPT_TC_FORMAT, fmt)
    e.set_option(FFB_OPT_TENSOR_CORE, 2)
    e.set_option(FFB_OPT_DEDUP_PAD, 0)
    b = g["batch"]
    e.encode(b["input"].reshape(b["input"].shape[0], b["input"].shape[1], -1), b["input_mask"], b["num_input"])
    lg = e.forced_prefix_logits(g["prefix"])
    ok, d = logits_close(lg, g["prefix_logits"], b64=g.get("prefix_logits64"))
    assert ok, f"forced-prefix logits differ by {d}"
    e.close()


def test_tensor_core_and_simt_paths_agree_on_a_real_batch():
    """8 wireframes at ours.yml size (M up to ~30k rows, auto mode picks the tensor-core path):
    identical tokens and logits within 1e-4 of the all-SIMT run."""
    from faceformer_b200 import synth
    cfg = OURS
    sd = synth.synth_state_dict(cfg, MODE_PARALLEL, 0, "diverse")
    batch = synth.synth_batch(cfg, MODE_PARALLEL, 8, seed=3, lo=40, hi=120)
    coords = torch.from_numpy(batch["input"]).cuda().flatten(2)
    mask, ni = torch.from_numpy(batch["input_mask"]).cuda(), torch.from_numpy(batch["num_input"]).cuda()
    out = []
    for mode in (0, 1):
        e = Engine(cfg, MODE_PARALLEL, 0)
        e.load_state_dict(sd)
        e.set_option(FFB_OPT_TENSOR_CORE, mode)
        from faceformer_b200.lib import FFB_OPT_PROFILE
        e.set_option(FFB_OPT_PROFILE, 1)
        pred, steps = e.forward_eval(coords, mask, ni)
        prof = e.profile_read()
        out.append((pred.cpu().numpy(), steps, e.get_last_logits().cpu().numpy(), prof))
        e.close()
    assert out[1][3]["linear_tc"]["launches"] > 0 and out[0][3]["linear_tc"]["launches"] == 0
    assert out[0][1] == out[1][1]
    assert np.array_equal(out[0][0], out[1][0]), f"{(out[0][0] != out[1][0]).sum()} token mismatches"
    ok, d = logits_close(out[1][2], out[0][2])
    assert ok, d


@pytest.mark.parametrize("attn_x", [1, 2, 3])
def test_tcgen05_attention_agrees_with_mma_sync_on_a_real_batch(attn_x):
    """Same batch decoded with the attention cores on tcgen05 (attn_x_kernel; bit 0 cross, bit 1 self) and on
    mma.sync (attn_h_kernel): identical tokens, logits within tolerance; ragged wireframes (different key counts, tiles
    straddling wireframes and sequences)."""
    from faceformer_b200 import synth
    from faceformer_b200.lib import FFB_OPT_ATTN_X
    cfg = OURS
    sd = synth.synth_state_dict(cfg, MODE_PARALLEL, 0, "diverse")
    batch = synth.synth_batch(cfg, MODE_PARALLEL, 9, seed=5, lo=24, hi=216)
    coords = torch.from_numpy(batch["input"]).cuda().flatten(2)
    mask, ni = torch.from_numpy(batch["input_mask"]).cuda(), torch.from_numpy(batch["num_input"]).cuda()
    out = []
    for mode in (0, attn_x):
        e = Engine(cfg, MODE_PARALLEL, 0)
        e.load_state_dict(sd)
        e.set_option(FFB_OPT_TENSOR_CORE, 2)
        e.set_option(FFB_OPT_ATTN_X, mode)
        pred, steps = e.forward_eval(coords, mask, ni)
        out.append((pred.cpu().numpy(), steps, e.get_last_logits().cpu().numpy(), e.fp16_fallbacks()))
        e.close()
    assert out[0][3] == 0 and out[1][3] == 0
    assert out[0][1] == out[1][1]
    assert np.array_equal(out[0][0], out[1][0]), f"{(out[0][0] != out[1][0]).sum()} token mismatches"
    ok, d = logits_close(out[1][2], out[0][2])
    assert ok, d


def test_encoder_modes_match_the_oracle_encoder():
    """A batch with > 2048 memory rows through the three encoder modes: fp32 SIMT, the tcgen05 fp16x2 pipeline (the throughput mode) and
    float64 (the default, enc64.cuh).  Encoder memory against the oracle's fp32 encoder: within 1e-4 each; the float64 encoder must sit
    in the middle (closer to both fp32-class evaluations than they are to each other is not required, but it must not be farther)."""
    from faceformer_b200 import synth
    from faceformer_b200.lib import FFB_OPT_ENCODER_PRECISION, FFB_OPT_ENCODER_TC, FFB_OPT_PROFILE
    from oracle import faceformer_oracle as orc
    cfg = OURS
    sd = synth.synth_state_dict(cfg, MODE_PARALLEL, 0, "diverse")
    batch = synth.synth_batch(cfg, MODE_PARALLEL, 20, seed=11, lo=60, hi=216)
    coords = torch.from_numpy(batch["input"]).cuda().flatten(2)
    mask, ni = torch.from_numpy(batch["input_mask"]).cuda(), torch.from_numpy(batch["num_input"]).cuda()
    mems = []
    for prec, enc_tc in ((0, 0), (0, 1), (2, 1)):
        e = Engine(cfg, MODE_PARALLEL, 0)
        e.load_state_dict(sd)
        e.set_option(FFB_OPT_ENCODER_PRECISION, prec)
        e.set_option(FFB_OPT_ENCODER_TC, enc_tc)
        e.set_option(FFB_OPT_PROFILE, 1)
        info = e.encode(coords, mask, ni)
        prof = e.profile_read()
        assert info["R"] >= 2048
        if prec == 0:
            assert (prof["linear_tc"]["launches"] > 0) == bool(enc_tc)
        else:       # float64 encoder; on a batch this large only the two cross-attention K / V projections run on the tensor cores
            assert prof["linear_tc"]["launches"] == 2
        mems.append(e.get_memory().cpu().numpy())
        assert e.fp16_fallbacks() == 0
        e.close()
    want = orc.encode(sd, cfg.to_dict(), MODE_PARALLEL, batch)[0].transpose(1, 0, 2)      # [N, L, E]
    vm = valid_rows_mask(batch, cfg)
    d = [float(np.max(np.abs(m[vm] - want[vm]))) for m in mems]
    assert max(d) <= LOGIT_TOL, d
    assert np.max(np.abs(mems[1][vm] - mems[0][vm])) <= LOGIT_TOL
    assert np.max(np.abs(mems[2][vm] - mems[0][vm])) <= LOGIT_TOL and np.max(np.abs(mems[2][vm] - mems[1][vm])) <= LOGIT_TOL


def test_more_wireframes_than_the_tcgen05_attention_groups():
    """260 small wireframes in one batch: more groups than attn_x_kernel's prefix table holds (255), so the cross-attention and the
    encoder self-attention take the mma.sync kernel while GEMMs and decoder self-attention stay on tcgen05.  Tokens equal to the
    all-SIMT run, logits within tolerance."""
    from faceformer_b200 import synth
    cfg = OURS
    sd = synth.synth_state_dict(cfg, MODE_PARALLEL, 0, "diverse")
    batch = synth.synth_batch(cfg, MODE_PARALLEL, 260, seed=17, lo=6, hi=14)
    coords = torch.from_numpy(batch["input"]).cuda().flatten(2)
    mask, ni = torch.from_numpy(batch["input_mask"]).cuda(), torch.from_numpy(batch["num_input"]).cuda()
    out = []
    for mode in (0, 1):
        e = Engine(cfg, MODE_PARALLEL, 0)
        e.load_state_dict(sd)
        e.set_option(FFB_OPT_TENSOR_CORE, mode)
        pred, steps = e.forward_eval(coords, mask, ni)
        out.append((pred.cpu().numpy(), steps, e.get_last_logits().cpu().numpy(), e.fp16_fallbacks(), e.batch_info()))
        e.close()
    assert out[1][4]["N"] == 260 and out[1][4]["R"] >= 2048 and out[1][3] == 0
    assert out[0][1] == out[1][1]
    assert np.array_equal(out[0][0], out[1][0]), f"{(out[0][0] != out[1][0]).sum()} token mismatches"
    ok, d = logits_close(out[1][2], out[0][2], tol=2e-4)
    assert ok, d


def test_fp16_overflow_falls_back_to_bf16x3():
    """FFN hidden activations beyond the fp16 range (linear1 x 65536, linear2 / 65536: the same function in exact
    arithmetic): the fp16x2 decode raises the overflow flag, is re-run in bf16x3 and still matches the SIMT path."""
    g = load_case("ours_parallel_small")
    sd = {k: v.copy() for k, v in g["sd"].items()}
    for k in sd:
        if k.startswith("decoder.") and k.endswith("linear1.weight") or k.startswith("decoder.") and k.endswith("linear1.bias"):
            sd[k] = sd[k] * np.float32(65536.0)
        if k.startswith("decoder.") and k.endswith("linear2.weight"):
            sd[k] = sd[k] / np.float32(65536.0)
    b = g["batch"]
    coords = torch.from_numpy(b["input"]).cuda().flatten(2)
    mask, ni = torch.from_numpy(b["input_mask"]).cuda(), torch.from_numpy(b["num_input"]).cuda()
    res = []
    for tc_mode in (0, 2):
        e = Engine(g["cfg"], g["mode"], 0)
        e.load_state_dict(sd)
        e.set_option(FFB_OPT_TENSOR_CORE, tc_mode)
        pred, steps = e.forward_eval(coords, mask, ni)
        res.append((pred.cpu().numpy(), steps, e.get_last_logits().cpu().numpy(), e.fp16_fallbacks()))
        e.close()
    assert res[0][3] == 0 and res[1][3] == 1                       # exactly one re-run, then the handle stays in bf16x3
    assert res[0][1] == res[1][1] and np.array_equal(res[0][0], res[1][0])
    assert np.array_equal(res[0][0], g["predict"])                  # power-of-two rescaling is exact: same tokens as the golden
    for r in res:                                                   # both runs against the reference's logits (noise-floor clause, util.py)
        ok, d = logits_close(r[2], g["last_logits"], b64=g.get("last_logits64"))
        assert ok, d


def test_async_decode_reports_fp16_overflow():
    """a fully asynchronous decode (device buffers, no steps_run) cannot re-run by itself; ffb_overflowed() must report
    the fp16-range overflow, switch the handle to bf16x3 and the re-run must then match the SIMT path."""
    g = load_case("ours_parallel_small")
    sd = {k: v.copy() for k, v in g["sd"].items()}
    for k in sd:
        if k.startswith("decoder.") and (k.endswith("linear1.weight") or k.endswith("linear1.bias")):
            sd[k] = sd[k] * np.float32(65536.0)
        if k.startswith("decoder.") and k.endswith("linear2.weight"):
            sd[k] = sd[k] / np.float32(65536.0)
    b = g["batch"]
    coords = torch.from_numpy(b["input"]).cuda().flatten(2)
    mask, ni = torch.from_numpy(b["input_mask"]).cuda(), torch.from_numpy(b["num_input"]).cuda()
    e = Engine(g["cfg"], g["mode"], 0)
    e.load_state_dict(sd)
    e.set_option(FFB_OPT_TENSOR_CORE, 2)
    pred, steps = e.forward_eval(coords, mask, ni, want_steps=False)
    assert steps is None
    assert e.overflowed() is True and e.fp16_fallbacks() == 1
    with pytest.raises(Exception):
        e.get_last_logits()                                       # the invalid decode is no longer readable
    pred, _ = e.forward_eval(coords, mask, ni, want_steps=False)   # now in bf16x3
    assert e.overflowed() is False and e.fp16_fallbacks() == 1
    assert np.array_equal(pred.cpu().numpy(), g["predict"])
    e.close()
    # a healthy asynchronous decode reports no overflow
    e = Engine(g["cfg"], g["mode"], 0)
    e.load_state_dict(g["sd"])
    e.set_option(FFB_OPT_TENSOR_CORE, 2)
    pred, _ = e.forward_eval(coords, mask, ni, want_steps=False)
    assert e.overflowed() is False and e.fp16_fallbacks() == 0
    assert np.array_equal(pred.cpu().numpy(), g["predict"])
    e.close()


def test_reload_and_option_change_invalidate_the_encoded_batch():
    """stale encoder state must not survive a weight reload or an option that feeds the per-batch plan."""
    from faceformer_b200.lib import FFB_OPT_ATTN_MMA, FFBError
    g = load_case("ours_parallel_small")
    b = g["batch"]
    coords = torch.from_numpy(b["input"]).cuda().flatten(2)
    mask, ni = torch.from_numpy(b["input_mask"]).cuda(), torch.from_numpy(b["num_input"]).cuda()
    e = Engine(g["cfg"], g["mode"], 0)
    e.load_state_dict(g["sd"])
    e.encode(coords, mask, ni)
    e.load_state_dict(g["sd"])
    with pytest.raises(FFBError):
        e.decode_greedy()
    e.encode(coords, mask, ni)
    e.set_option(FFB_OPT_ATTN_MMA, 1)
    with pytest.raises(FFBError):
        e.decode_greedy()
    pred, steps = e.forward_eval(coords, mask, ni)
    assert np.array_equal(pred.cpu().numpy(), g["predict"]) and steps == g["steps"]
    e.close()


@pytest.mark.parametrize("name", ["mid_parallel_trained", "mid_parallel_trained_b"])
@pytest.mark.parametrize("fmt", [2, 3])
def test_trained_e512_checkpoint_through_forced_tcgen05(name, fmt):
    """non-degenerate (trained) weights on the tensor path.  An E = 512 / H = 8 checkpoint trained with the reference's
    own forward_train (oracle/train_fixture.py --cfg mid; 27 distinct tokens in the goldens), 16 / 24 wireframes, EVERY decode step forced
    through tc::gemm_kernel / ax::attn_x_kernel, against the unmodified reference's output: tokens exact, logits within 1e-4."""
    g = load_case(name)
    e = Engine(g["cfg"], g["mode"], 0)
    e.load_state_dict(g["sd"])
    e.set_option(FFB_OPT_TC_FORMAT, fmt)
    e.set_option(FFB_OPT_TENSOR_CORE, 2)
    b = g["batch"]
    coords = torch.from_numpy(b["input"]).cuda().flatten(2)
    mask, ni = torch.from_numpy(b["input_mask"]).cuda(), torch.from_numpy(b["num_input"]).cuda()
    from faceformer_b200.lib import FFB_OPT_PROFILE
    e.set_option(FFB_OPT_PROFILE, 1)
    pred, steps = e.forward_eval(coords, mask, ni)
    prof = e.profile_read()
    e.set_option(FFB_OPT_PROFILE, 0)
    assert prof["linear_tc"]["launches"] > 0 and e.fp16_fallbacks() == 0
    assert steps == g["steps"] and np.array_equal(pred.cpu().numpy(), g["predict"])
    assert len(np.unique(g["predict"])) >= 20                                          # the fixture is not degenerate
    ok, d = logits_close(e.get_last_logits().cpu().numpy(), g["last_logits"], b64=g.get("last_logits64"))
    assert ok, d
    vm = valid_rows_mask(b, g["cfg"])
    assert np.max(np.abs(e.get_memory().cpu().numpy()[vm] - g["memory"][vm])) <= LOGIT_TOL
    e.close()


@pytest.mark.parametrize("name", ["ours_parallel_small", "mid_parallel_trained", "perspective_small"])
def test_layer0_cache_is_exact(name):
    """FFB_OPT_L0_CACHE: decoder layer 0 projects only the NEW prefix position and takes q / k / v of the earlier ones from a cache.  Their
    inputs cannot change between steps, every row goes through the same LayerNorm and the same K-loop, so tokens AND logits must be
    bit-identical to the run that recomputes them (and equal to the reference golden)."""
    from faceformer_b200.lib import FFB_OPT_L0_CACHE
    g = load_case(name)
    b = g["batch"]
    coords = torch.from_numpy(b["input"]).cuda().flatten(2)
    mask, ni = torch.from_numpy(b["input_mask"]).cuda(), torch.from_numpy(b["num_input"]).cuda()
    res = []
    for cache in (0, 1):
        e = Engine(g["cfg"], g["mode"], 0)
        e.load_state_dict(g["sd"])
        e.set_option(FFB_OPT_L0_CACHE, cache)
        pred, steps = e.forward_eval(coords, mask, ni)
        res.append((pred.cpu().numpy(), steps, e.get_last_logits().cpu().numpy(), e.kernel_launches()))
        e.close()
    assert res[0][1] == res[1][1] == g["steps"]
    assert np.array_equal(res[0][0], g["predict"]) and np.array_equal(res[1][0], g["predict"])
    assert np.array_equal(res[0][2], res[1][2])
