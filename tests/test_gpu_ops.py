"""Op-level parity: each CUDA kernel of the path, called through the C ABI, against the oracle's
numpy primitive on the same seeded inputs (level i)."""
import numpy as np
import pytest
import torch

from faceformer_b200.config import MODE_PARALLEL, OURS, TINY
from faceformer_b200.engine import Engine
from oracle import faceformer_oracle as orc

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def eng():
    e = Engine(OURS, MODE_PARALLEL, 0)
    yield e
    e.close()


def _t(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


@pytest.mark.parametrize("M,N,K", [(1, 128, 128), (37, 512, 512), (300, 1536, 512), (129, 512, 1024),
                                   (1000, 512, 100), (257, 384, 128), (5000, 1024, 512)])
@pytest.mark.parametrize("variant", ["plain", "bias_relu", "bias_res", "pos"])
def test_linear(eng, M, N, K, variant):
    rng = np.random.default_rng(M * 7 + N + K)
    A = rng.normal(size=(M, K)).astype(np.float32)
    W = (rng.normal(size=(N, K)) / np.sqrt(K)).astype(np.float32)
    b = rng.normal(size=N).astype(np.float32)
    R = rng.normal(size=(M, N)).astype(np.float32)
    if variant == "plain":
        got = eng.op_linear(_t(A), _t(W))
        want = orc.linear(A, W)
    elif variant == "bias_relu":
        got = eng.op_linear(_t(A), _t(W), bias=_t(b), relu=True)
        want = np.maximum(orc.linear(A, W, b), 0)
    elif variant == "bias_res":
        got = eng.op_linear(_t(A), _t(W), bias=_t(b), R=_t(R))
        want = R + orc.linear(A, W, b)
    else:
        if N % 128 != 0 or N < 256:
            pytest.skip("pos variant needs >= 2 column tiles")
        P = 7
        pos = rng.normal(size=(P, K)).astype(np.float32)
        pos_cols = (N // 256) * 128                       # first tiles get +pos, the rest do not
        got = eng.op_linear(_t(A), _t(W), bias=_t(b), pos=_t(pos), pos_mod=P, pos_cols=pos_cols)
        Ap = A + pos[np.arange(M) % P]
        want = np.concatenate([orc.linear(Ap, W[:pos_cols], b[:pos_cols]), orc.linear(A, W[pos_cols:], b[pos_cols:])], 1)
    got = got.cpu().numpy()
    ref64 = None
    assert got.shape == want.shape
    err = np.max(np.abs(got - want))
    assert err <= 2e-5 * max(1.0, float(np.abs(want).max())), err


@pytest.mark.parametrize("M,E", [(1, 128), (33, 512), (1000, 512), (17, 1024)])
def test_layernorm(eng, M, E):
    rng = np.random.default_rng(M + E)
    x = (rng.normal(size=(M, E)) * 3 + 1).astype(np.float32)
    g = rng.normal(size=E).astype(np.float32)
    b = rng.normal(size=E).astype(np.float32)
    got = eng.op_layernorm(_t(x), _t(g), _t(b)).cpu().numpy()
    want = orc.layer_norm(x, g, b)
    assert np.max(np.abs(got - want)) <= 1e-5


def _attn_ref(q, k, v, G, nq, nk, H):
    out = np.zeros((G * nq, H * 64), np.float32)
    for g in range(G):
        for h in range(H):
            qq = q[g * nq:(g + 1) * nq, h * 64:(h + 1) * 64] * np.float32(0.125)
            kk = k[g * nk:(g + 1) * nk, h * 64:(h + 1) * 64]
            vv = v[g * nk:(g + 1) * nk, h * 64:(h + 1) * 64]
            p = orc.softmax_lastdim(qq @ kk.T)
            out[g * nq:(g + 1) * nq, h * 64:(h + 1) * 64] = p @ vv
    return out


@pytest.mark.parametrize("kind", [0, 1, 2, 3])
@pytest.mark.parametrize("G,nq,nk", [(3, 1, 1), (5, 7, 7), (4, 36, 36), (2, 37, 37), (2, 70, 70), (3, 64, 220),
                                     (2, 130, 65), (1, 258, 258), (2, 5, 516)])
def test_attention(eng, kind, G, nq, nk):
    H = OURS.num_head
    rng = np.random.default_rng(G * 100 + nq + nk)
    q = (rng.normal(size=(G * nq, H * 64)) * 1.5).astype(np.float32)
    kv = (rng.normal(size=(G * nk, 2 * H * 64)) * 1.5).astype(np.float32)     # k and v interleaved in one buffer (ld = 2E)
    kvt = _t(kv)
    got = eng.op_attention(kind, _t(q), kvt[:, :H * 64], kvt[:, H * 64:], G, nq, nk).cpu().numpy()
    want = _attn_ref(q, kv[:, :H * 64], kv[:, H * 64:], G, nq, nk, H)
    assert np.max(np.abs(got - want)) <= 2e-5


@pytest.mark.parametrize("G,nq,nk", [(3, 1, 1), (5, 7, 7), (4, 36, 36), (2, 37, 37), (2, 70, 70), (3, 64, 220),
                                     (2, 130, 65), (1, 258, 258), (2, 5, 516), (40, 1, 36)])
def test_attention_half_inputs(eng, G, nq, nk):
    """attn_h_kernel: q/k/v arrive as fp16x2 splits (the hook splits them), staged with cp.async, fragments from ldmatrix."""
    H = OURS.num_head
    rng = np.random.default_rng(G * 100 + nq + nk + 1)
    q = (rng.normal(size=(G * nq, H * 64)) * 1.5).astype(np.float32)
    k = (rng.normal(size=(G * nk, H * 64)) * 1.5).astype(np.float32)
    v = (rng.normal(size=(G * nk, H * 64)) * 1.5).astype(np.float32)
    got = eng.op_attention(4, _t(q), _t(k), _t(v), G, nq, nk).cpu().numpy()
    want = _attn_ref(q, k, v, G, nq, nk, H)
    assert np.max(np.abs(got - want)) <= 2e-5


CROSS_CASES = [(3, 1, 1), (5, 7, 7), (4, 36, 36), (2, 37, 37), (2, 130, 65), (3, 64, 220), (1, 258, 256), (6, 300, 33), (40, 1, 36),
               (9, 513, 100), (160, 140, 224)]


@pytest.mark.parametrize("G,nq,nk", CROSS_CASES)
def test_attention_tcgen05_cross(eng, G, nq, nk, kind=5):
    """attn_x_kernel (tcgen05 + TMEM + TMA), CROSS mode: S = Q K^T and O = P V as fp16x2 split products, softmax in fp32,
    <= 256 keys, V rows consumed through an MN-major descriptor.  The last case has more work items than SMs x 8, so persistent
    CTAs walk several (group, head) pairs and tiles."""
    H = OURS.num_head
    rng = np.random.default_rng(G * 100 + nq + nk + 2)
    q = (rng.normal(size=(G * nq, H * 64)) * 1.5).astype(np.float32)
    k = (rng.normal(size=(G * nk, H * 64)) * 1.5).astype(np.float32)
    v = (rng.normal(size=(G * nk, H * 64)) * 1.5).astype(np.float32)
    got = eng.op_attention(kind, _t(q), _t(k), _t(v), G, nq, nk).cpu().numpy()
    want = _attn_ref(q, k, v, G, nq, nk, H)
    assert np.max(np.abs(got - want)) <= 2e-5


@pytest.mark.parametrize("G,P", [(1, 1), (3, 1), (300, 1), (7, 2), (5, 7), (40, 16), (9, 31), (10, 32), (11, 33), (100, 36), (7, 37),
                                 (3, 64), (5, 65), (2, 128), (3000, 36), (2500, 5)])
def test_attention_tcgen05_self(eng, G, P):
    """attn_x_kernel, SELF mode: tiles of floor(128 / P) whole sequences, block-diagonal softmax (decoder self-attention, no mask)."""
    H = OURS.num_head
    rng = np.random.default_rng(G * 100 + P + 3)
    q = (rng.normal(size=(G * P, H * 64)) * 1.5).astype(np.float32)
    k = (rng.normal(size=(G * P, H * 64)) * 1.5).astype(np.float32)
    v = (rng.normal(size=(G * P, H * 64)) * 1.5).astype(np.float32)
    got = eng.op_attention(6, _t(q), _t(k), _t(v), G, P, P).cpu().numpy()
    want = _attn_ref(q, k, v, G, P, P, H)
    assert np.max(np.abs(got - want)) <= 2e-5


@pytest.mark.timeout(120)
@pytest.mark.parametrize("G,n", [(1, 1), (3, 7), (2, 63), (3, 64), (2, 65), (5, 129), (2, 300), (3, 516), (1, 2052), (40, 220), (2, 1000)])
def test_attention_tcgen05_long(eng, G, n):
    """attn_long_kernel (attn_l.cuh): encoder self-attention with any number of keys -- K / V stream through shared memory in 64-key
    blocks, exact two-pass softmax, O drained every 8 blocks.  2052 keys = the 2048-edge wireframes of BASELINE.json configs[4]."""
    H = OURS.num_head
    rng = np.random.default_rng(G * 100 + n + 4)
    q = (rng.normal(size=(G * n, H * 64)) * 1.5).astype(np.float32)
    k = (rng.normal(size=(G * n, H * 64)) * 1.5).astype(np.float32)
    v = (rng.normal(size=(G * n, H * 64)) * 1.5).astype(np.float32)
    got = eng.op_attention(7, _t(q), _t(k), _t(v), G, n, n).cpu().numpy()
    want = _attn_ref(q, k, v, G, n, n, H)
    assert np.max(np.abs(got - want)) <= 2e-5
