// attn_f16.cuh -- attention core on the tensor pipe, fp16x2 split precision (mma.sync m16n8k16, fp32 accumulate).
//
//   O = softmax(Q K^T / 8) V   per (group, head), head dim 64, any number of keys (online softmax, 64-key tiles)
//
// Same contract as attn_mma.cuh (3xTF32), half the tensor work: every fp32 operand x is carried as two halves
// x = h + l (h = RN_f16(x), l = RN_f16(x - h), exact to 2^-22 in fp16's normal range) and every product as
// h*h + (l*h + h*l) -- 3 MMAs per 16-wide k-step instead of 6 per two 8-wide ones.  The split is done ONCE per tile
// while staging into shared memory (the 1/8 score scale is applied to the fp32 scores, exact).  |x| > 65504 raises the shared overflow flag
// (the host then re-runs the decode with the TF32 attention kernel and the bf16x3 GEMM).
//
// Shared-memory layouts are permuted so that every fragment is two conflict-free LDS.128:
//   * head dim d of Q/K rows is stored at column 16*t + 4*s + 2*j + e  with  d = 16*s + 8*j + 2*t + e
//     (s = k-step, t = thread-in-quad, j = low/high k-half, e = element of the pair): thread t owns 16 contiguous halves;
//   * V is staged row-major ([key][head dim]); ldmatrix.x4.trans delivers the (key-pair, head-dim) B fragments of P.V
//     (row stride 144 B: conflict-free), and the S accumulators of key blocks 2s, 2s+1 ARE the A fragment of k-step s;
//   * two lanes of a quad swap one pair per tile pair at the end, so every lane stores 4 contiguous outputs.
#pragma once
#include "kernels.cuh"

namespace ffb {

constexpr int AF_BQ = 64, AF_BK = 64, AF_S = 72;                        // row stride in halves (144 B: conflict-free LDS.128)
constexpr int AF_TILE = 64 * AF_S;                                      // halves per tile
constexpr int AF_SMEM_BYTES = 6 * AF_TILE * 2;                          // Qh Ql Kh Kl Vth Vtl = 55,296 B

__device__ __forceinline__ void mma_f16(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile(
        "mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
        : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ int af_perm(int i) {                         // i in [0,64): index within a 64-wide contraction block
    return 16 * ((i & 7) >> 1) + 4 * (i >> 4) + 2 * ((i >> 3) & 1) + (i & 1);
}
__device__ __forceinline__ uint32_t pack_h2(__half a, __half b) {
    return (uint32_t)__half_as_ushort(a) | ((uint32_t)__half_as_ushort(b) << 16);
}
// split two fp32 values into packed hi / lo half2 words (packed conversions: one F2FP per pair)
__device__ __forceinline__ void split_pair(float x, float y, uint32_t& hi, uint32_t& lo) {
    const __half2 h = __floats2half2_rn(x, y);
    const float2 hf = __half22float2(h);
    const __half2 l = __floats2half2_rn(x - hf.x, y - hf.y);
    hi = *reinterpret_cast<const uint32_t*>(&h);
    lo = *reinterpret_cast<const uint32_t*>(&l);
}
// exp(x) for x <= 0 on the SFU: ex2.approx(x * log2 e), relative error ~2^-22 (the softmax weights are split to 22 bits anyway)
__device__ __forceinline__ float fast_exp(float x) { return exp2f(x * 1.4426950408889634f); }

__global__ void __launch_bounds__(128, 3) attn_f16_kernel(const float* __restrict__ Q, int ldq,
                                                          const float* __restrict__ K, const float* __restrict__ V, int ldk,
                                                          float* __restrict__ O, int ldo, uint16_t* __restrict__ Os,
                                                          long long os_stride, const AttnGroups g, const int* stop) {
    FFB_STOP_CHECK(stop);
    extern __shared__ __align__(16) uint16_t smem_h[];
    uint16_t* Qh = smem_h;              uint16_t* Ql = smem_h + AF_TILE;
    uint16_t* Kh = smem_h + 2 * AF_TILE; uint16_t* Kl = smem_h + 3 * AF_TILE;
    uint16_t* Vh = smem_h + 4 * AF_TILE; uint16_t* Vl = smem_h + 5 * AF_TILE;     // [key][head dim]

    long long q0, k0, o0; int nq, nk;
    attn_group(g, blockIdx.x, q0, nq, k0, nk, o0);
    const int head = blockIdx.y;
    const int qt0 = blockIdx.z * AF_BQ;
    if (qt0 >= nq) return;
    const int nqt = min(AF_BQ, nq - qt0);
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    const int gq = lane >> 2, t = lane & 3;
    bool ovf = false;

    // ---- stage Q as hi/lo halves (the 1/8 score scale is applied after the MMAs: exact), head dim permuted; rows beyond nqt are zero (only up to the last active warp) ----
    const int q_rows = min(AF_BQ, (nqt + 15) & ~15);
    for (int idx = tid; idx < q_rows * 16; idx += 128) {
        const int r = idx >> 4, d4 = idx & 15;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (r < nqt) {
            v = *reinterpret_cast<const float4*>(Q + (size_t)(q0 + qt0 + r) * ldq + head * 64 + d4 * 4);
        }
        ovf |= !(fabsf(v.x) <= 65504.f) | !(fabsf(v.y) <= 65504.f) | !(fabsf(v.z) <= 65504.f) | !(fabsf(v.w) <= 65504.f);
        uint32_t h0, l0, h1, l1;
        split_pair(v.x, v.y, h0, l0); split_pair(v.z, v.w, h1, l1);
        const int c0 = af_perm(d4 * 4), c1 = af_perm(d4 * 4 + 2);        // pairs (d, d+1) stay adjacent under the permutation
        *reinterpret_cast<uint32_t*>(Qh + r * AF_S + c0) = h0; *reinterpret_cast<uint32_t*>(Ql + r * AF_S + c0) = l0;
        *reinterpret_cast<uint32_t*>(Qh + r * AF_S + c1) = h1; *reinterpret_cast<uint32_t*>(Ql + r * AF_S + c1) = l1;
    }
    __syncthreads();

    const bool warp_active = (w * 16) < nqt;
    uint32_t qh[4][4], ql[4][4];                                         // A fragments of the 4 k-steps (hi, lo)
    if (warp_active) {
#pragma unroll
        for (int half = 0; half < 2; ++half) {                           // uint4 #half covers k-steps 2*half, 2*half+1
            const uint4 h0 = *reinterpret_cast<const uint4*>(Qh + (w * 16 + gq) * AF_S + 16 * t + 8 * half);
            const uint4 h1 = *reinterpret_cast<const uint4*>(Qh + (w * 16 + gq + 8) * AF_S + 16 * t + 8 * half);
            const uint4 l0 = *reinterpret_cast<const uint4*>(Ql + (w * 16 + gq) * AF_S + 16 * t + 8 * half);
            const uint4 l1 = *reinterpret_cast<const uint4*>(Ql + (w * 16 + gq + 8) * AF_S + 16 * t + 8 * half);
            // a0 = (row g, k 2t..) a1 = (row g+8, k 2t..) a2 = (row g, k 2t+8..) a3 = (row g+8, k 2t+8..)
            qh[2 * half][0] = h0.x; qh[2 * half][1] = h1.x; qh[2 * half][2] = h0.y; qh[2 * half][3] = h1.y;
            qh[2 * half + 1][0] = h0.z; qh[2 * half + 1][1] = h1.z; qh[2 * half + 1][2] = h0.w; qh[2 * half + 1][3] = h1.w;
            ql[2 * half][0] = l0.x; ql[2 * half][1] = l1.x; ql[2 * half][2] = l0.y; ql[2 * half][3] = l1.y;
            ql[2 * half + 1][0] = l0.z; ql[2 * half + 1][1] = l1.z; ql[2 * half + 1][2] = l0.w; ql[2 * half + 1][3] = l1.w;
        }
    }

    float m0 = -INFINITY, m1 = -INFINITY, l0s = 0.f, l1s = 0.f;
    float o[8][4];
#pragma unroll
    for (int u = 0; u < 8; ++u) { o[u][0] = o[u][1] = o[u][2] = o[u][3] = 0.f; }

    for (int kt = 0; kt < nk; kt += AF_BK) {
        const int nkt = min(AF_BK, nk - kt);
        const int k_rows = (nkt + 15) & ~15;                             // staged (zero-padded) keys: whole 16-key k-steps
        __syncthreads();                                                 // previous K/V tile fully consumed
        for (int idx = tid; idx < k_rows * 16; idx += 128) {
            const int r = idx >> 4, d4 = idx & 15;
            float4 kv = make_float4(0.f, 0.f, 0.f, 0.f), vv = kv;
            if (r < nkt) {
                const size_t off = (size_t)(k0 + kt + r) * ldk + head * 64 + d4 * 4;
                kv = *reinterpret_cast<const float4*>(K + off);
                vv = *reinterpret_cast<const float4*>(V + off);
            }
            ovf |= !(fabsf(kv.x) <= 65504.f) | !(fabsf(kv.y) <= 65504.f) | !(fabsf(kv.z) <= 65504.f) | !(fabsf(kv.w) <= 65504.f) |
                   !(fabsf(vv.x) <= 65504.f) | !(fabsf(vv.y) <= 65504.f) | !(fabsf(vv.z) <= 65504.f) | !(fabsf(vv.w) <= 65504.f);
            uint32_t h0, l0, h1, l1;
            split_pair(kv.x, kv.y, h0, l0); split_pair(kv.z, kv.w, h1, l1);
            const int c0 = af_perm(d4 * 4), c1 = af_perm(d4 * 4 + 2);
            *reinterpret_cast<uint32_t*>(Kh + r * AF_S + c0) = h0; *reinterpret_cast<uint32_t*>(Kl + r * AF_S + c0) = l0;
            *reinterpret_cast<uint32_t*>(Kh + r * AF_S + c1) = h1; *reinterpret_cast<uint32_t*>(Kl + r * AF_S + c1) = l1;
            split_pair(vv.x, vv.y, h0, l0); split_pair(vv.z, vv.w, h1, l1);          // V row-major, natural head-dim order
            *reinterpret_cast<uint2*>(Vh + r * AF_S + d4 * 4) = make_uint2(h0, h1);
            *reinterpret_cast<uint2*>(Vl + r * AF_S + d4 * 4) = make_uint2(l0, l1);
        }
        __syncthreads();
        if (!warp_active) continue;
        const int jmax = (nkt + 7) >> 3;                                 // 8-key blocks that contain valid keys
        const int smax = (nkt + 15) >> 4;                                // 16-key k-steps of the P.V product

        // ---- S = Q K^T: s[j] = keys 8j..8j+7 in C layout ----
        float s[8][4];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            if (j < jmax) {
                float sm[4] = {0.f, 0.f, 0.f, 0.f}, sc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
                for (int half = 0; half < 2; ++half) {
                    const uint4 kh = *reinterpret_cast<const uint4*>(Kh + (8 * j + gq) * AF_S + 16 * t + 8 * half);
                    const uint4 kl = *reinterpret_cast<const uint4*>(Kl + (8 * j + gq) * AF_S + 16 * t + 8 * half);
                    mma_f16(sc, ql[2 * half], kh.x, kh.y); mma_f16(sc, qh[2 * half], kl.x, kl.y); mma_f16(sm, qh[2 * half], kh.x, kh.y);
                    mma_f16(sc, ql[2 * half + 1], kh.z, kh.w); mma_f16(sc, qh[2 * half + 1], kl.z, kl.w); mma_f16(sm, qh[2 * half + 1], kh.z, kh.w);
                }
                const int key = kt + 8 * j + 2 * t;
                s[j][0] = (key < nk) ? (sm[0] + sc[0]) * 0.125f : -INFINITY; s[j][1] = (key + 1 < nk) ? (sm[1] + sc[1]) * 0.125f : -INFINITY;
                s[j][2] = (key < nk) ? (sm[2] + sc[2]) * 0.125f : -INFINITY; s[j][3] = (key + 1 < nk) ? (sm[3] + sc[3]) * 0.125f : -INFINITY;
            } else {
                s[j][0] = s[j][1] = s[j][2] = s[j][3] = -INFINITY;
            }
        }
        // ---- online softmax ----
        float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
        for (int j = 0; j < 8; ++j) { mx0 = fmaxf(mx0, fmaxf(s[j][0], s[j][1])); mx1 = fmaxf(mx1, fmaxf(s[j][2], s[j][3])); }
        mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1)); mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2));
        mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1)); mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
        const float mn0 = fmaxf(m0, mx0), mn1 = fmaxf(m1, mx1);
        const float corr0 = expf(m0 - mn0), corr1 = expf(m1 - mn1);
        float ps0 = 0.f, ps1 = 0.f;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            s[j][0] = expf(s[j][0] - mn0); s[j][1] = expf(s[j][1] - mn0);
            s[j][2] = expf(s[j][2] - mn1); s[j][3] = expf(s[j][3] - mn1);
            ps0 += s[j][0] + s[j][1]; ps1 += s[j][2] + s[j][3];
        }
        ps0 += __shfl_xor_sync(0xffffffffu, ps0, 1); ps0 += __shfl_xor_sync(0xffffffffu, ps0, 2);
        ps1 += __shfl_xor_sync(0xffffffffu, ps1, 1); ps1 += __shfl_xor_sync(0xffffffffu, ps1, 2);
        l0s = l0s * corr0 + ps0; l1s = l1s * corr1 + ps1;
        m0 = mn0; m1 = mn1;

        // ---- O_tile = P V from zero, then O = O * corr + O_tile in fp32 ----
        float om[8][4];
#pragma unroll
        for (int u = 0; u < 8; ++u) { om[u][0] = om[u][1] = om[u][2] = om[u][3] = 0.f; }
        // ldmatrix.x4.trans lane addressing: matrices (m0,m1) = keys 16ks + {0..7, 8..15} of head-dim block 8*(2up), (m2,m3) = same keys of
        // block 8*(2up+1); lane l supplies row (l & 7) of matrix (l >> 3).  Transposed, matrix m0/m1 are (b0, b1) of tile 2up, m2/m3 of tile 2up+1.
        const int lm_row = (lane & 7) + ((lane >> 3) & 1) * 8, lm_col = (lane >> 4) * 8;
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {
            if (ks < smax) {
                // A fragment = S accumulators of key blocks 2ks, 2ks+1.  Probabilities are mostly << 1/8, where the low half of the
                // split would be subnormal in fp16: split 4096*p instead (exact power of two, undone on O_tile below).
                uint32_t pa[4], pl[4];
                split_pair(s[2 * ks][0] * 4096.f, s[2 * ks][1] * 4096.f, pa[0], pl[0]);
                split_pair(s[2 * ks][2] * 4096.f, s[2 * ks][3] * 4096.f, pa[1], pl[1]);
                split_pair(s[2 * ks + 1][0] * 4096.f, s[2 * ks + 1][1] * 4096.f, pa[2], pl[2]);
                split_pair(s[2 * ks + 1][2] * 4096.f, s[2 * ks + 1][3] * 4096.f, pa[3], pl[3]);
                const uint32_t vh_addr = (uint32_t)__cvta_generic_to_shared(Vh + (16 * ks + lm_row) * AF_S + lm_col);
                const uint32_t vl_addr = (uint32_t)__cvta_generic_to_shared(Vl + (16 * ks + lm_row) * AF_S + lm_col);
#pragma unroll
                for (int up = 0; up < 4; ++up) {                         // output tiles 2up, 2up+1 (head dims 16up .. 16up+15)
                    uint32_t h0, h1, h2, h3, l0, l1, l2, l3;
                    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
                                 : "=r"(h0), "=r"(h1), "=r"(h2), "=r"(h3) : "r"(vh_addr + up * 32));
                    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
                                 : "=r"(l0), "=r"(l1), "=r"(l2), "=r"(l3) : "r"(vl_addr + up * 32));
                    mma_f16(om[2 * up], pl, h0, h1); mma_f16(om[2 * up], pa, l0, l1); mma_f16(om[2 * up], pa, h0, h1);
                    mma_f16(om[2 * up + 1], pl, h2, h3); mma_f16(om[2 * up + 1], pa, l2, l3); mma_f16(om[2 * up + 1], pa, h2, h3);
                }
            }
        }
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            constexpr float kInvP = 1.0f / 4096.f;
            o[u][0] = o[u][0] * corr0 + om[u][0] * kInvP; o[u][1] = o[u][1] * corr0 + om[u][1] * kInvP;
            o[u][2] = o[u][2] * corr1 + om[u][2] * kInvP; o[u][3] = o[u][3] * corr1 + om[u][3] * kInvP;
        }
    }
    if (ovf && g.overflow) *g.overflow = 1;
    if (!warp_active) return;

    // C layout: tile u holds head dims 8u + 2t + {0,1} of rows gq (o[u][0..1]) and gq+8 (o[u][2..3]).  Lanes t and t^1 swap
    // one pair per tile pair so that every lane owns 4 contiguous head dims (16-byte stores / 4-wide operand splits).
    const float inv0 = 1.0f / l0s, inv1 = 1.0f / l1s;
    const bool odd = (t & 1) != 0;
#pragma unroll
    for (int half = 0; half < 2; ++half) {
        const int r = w * 16 + gq + half * 8;
        const float inv = half ? inv1 : inv0;
        const int e = half * 2;
#pragma unroll
        for (int c = 0; c < 4; ++c) {                                    // tile pair (2c, 2c+1)
            const float a0 = o[2 * c][e] * inv, a1 = o[2 * c][e + 1] * inv, b0 = o[2 * c + 1][e] * inv, b1 = o[2 * c + 1][e + 1] * inv;
            const float s0 = odd ? a0 : b0, s1 = odd ? a1 : b1;          // odd lanes give away their tile-2c pair, even lanes their tile-(2c+1) pair
            const float r0 = __shfl_xor_sync(0xffffffffu, s0, 1), r1 = __shfl_xor_sync(0xffffffffu, s1, 1);
            const float4 v = odd ? make_float4(r0, r1, b0, b1) : make_float4(a0, a1, r0, r1);
            const int d = odd ? (8 * (2 * c + 1) + 2 * (t - 1)) : (8 * (2 * c) + 2 * t);
            if (r < nqt) {
                const size_t off = (size_t)(o0 + qt0 + r) * ldo + head * 64 + d;
                if (Os == nullptr) *reinterpret_cast<float4*>(O + off) = v;
                else store_split4(Os + off, os_stride, v, g.split_fmt, g.overflow);
            }
        }
    }
}

}  // namespace ffb
