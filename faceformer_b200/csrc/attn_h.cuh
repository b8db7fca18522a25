// attn_h.cuh -- attention core whose inputs ARRIVE as fp16x2 splits (the tensor-core GEMM epilogue writes q, k, v that way;
// the cross-attention K/V cache is split once per wireframe at encode time).
//
// Same arithmetic as attn_f16.cuh (hi*hi + lo*hi + hi*lo on mma.sync m16n8k16, fp32 accumulate, online softmax, P scaled
// by 4096 before its split, O_tile merged in fp32), but with ZERO operand-formatting work in the kernel:
//   * staging is a straight 16-byte cp.async copy of the hi and lo rows into natural row-major tiles (zero-filled past the end);
//   * all fragments come from ldmatrix (A of Q K^T: x4; B of Q K^T: x4 covering two key blocks; B of P V: x4.trans).
// Row stride 72 halves (144 B) makes every ldmatrix / cp.async conflict-free.
#pragma once
#include "attn_f16.cuh"

namespace ffb {

struct AttnHalfIn {
    const uint16_t* Qh; long long q_split; int ldq;          // lo part at Qh + q_split; row stride ldq halves
    const uint16_t* Kh; long long k_split;                   // keys
    const uint16_t* Vh; long long v_split; int ldk;          // values (same row stride as the keys)
};

__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, bool valid) {
    const int sz = valid ? 16 : 0;                           // src-size 0: nothing is read, 16 zero bytes are written
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(sz) : "memory");
}
__device__ __forceinline__ void ldsm_x4(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}

__device__ __forceinline__ float ex2_approx(float x) {          // 2^x on the SFU, one instruction; 2^-inf = +0
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

// Launch: blockDim.x = 32 * (number of 16-row query slabs of the largest tile, <= 4); dynamic smem = (2*qr_cap + 4*kr_cap) * 144 B
// where qr_cap / kr_cap = rows per Q / K,V tile buffer (multiples of 16, <= 64): short prefixes (decoder self-attention early in the
// decode) get small CTAs and many of them per SM.
__global__ void __launch_bounds__(128, 3) attn_h_kernel(const AttnHalfIn in, float* __restrict__ O, int ldo, uint16_t* __restrict__ Os,
                                                     long long os_stride, const AttnGroups g, int qr_cap, int kr_cap, const int* stop) {
    FFB_PDL_SYNC();
    FFB_STOP_CHECK(stop);
    extern __shared__ __align__(16) uint16_t smem_h[];
    const uint32_t qbytes = (uint32_t)qr_cap * AF_S * 2u, kbytes = (uint32_t)kr_cap * AF_S * 2u;
    const uint32_t sQh = (uint32_t)__cvta_generic_to_shared(smem_h), sQl = sQh + qbytes;
    const uint32_t sKh = sQl + qbytes, sKl = sKh + kbytes, sVh = sKl + kbytes, sVl = sVh + kbytes;

    long long q0, k0, o0; int nq, nk;
    attn_group(g, blockIdx.x, q0, nq, k0, nk, o0);
    const int head = blockIdx.y;
    const int qt0 = blockIdx.z * AF_BQ;
    if (qt0 >= nq) return;
    const int nqt = min(AF_BQ, nq - qt0);
    const int tid = threadIdx.x, nthr = blockDim.x, lane = tid & 31, w = tid >> 5;
    const int gq = lane >> 2, t = lane & 3;

    // ---- stage Q (hi, lo): 8 chunks of 16 B per row and part ----
    const int q_rows = min(qr_cap, (nqt + 15) & ~15);
    for (int idx = tid; idx < q_rows * 16; idx += nthr) {
        const int r = idx >> 4, c = idx & 7, part = (idx >> 3) & 1;
        const uint16_t* src = in.Qh + (part ? in.q_split : 0) + (size_t)(q0 + qt0 + min(r, nqt - 1)) * in.ldq + head * 64 + c * 8;
        cp_async16((part ? sQl : sQh) + (uint32_t)(r * AF_S + c * 8) * 2u, src, r < nqt);
    }
    asm volatile("cp.async.commit_group;" ::: "memory");

    const bool warp_active = (w * 16) < nqt;
    const int lm_row = (lane & 7) + ((lane >> 3) & 1) * 8, lm_col = (lane >> 4) * 8;       // A-operand / V (trans) lane addressing
    const int lk_row = (lane & 7) + (lane >> 4) * 8, lk_col = ((lane >> 3) & 1) * 8;       // K (B of Q K^T) lane addressing

    // Softmax state in the log2 domain, with the 4096x probability scale folded in:
    //   s' = s * log2(e)/8,  p' = 2^(s' - m' + 12) = 4096 * exp(s - m),  l' = sum p',  o' = sum p' v  ->  out = o' / l'
    float m0 = -INFINITY, m1 = -INFINITY, l0s = 0.f, l1s = 0.f;
    float o[8][4];
#pragma unroll
    for (int u = 0; u < 8; ++u) { o[u][0] = o[u][1] = o[u][2] = o[u][3] = 0.f; }
    uint32_t qh[4][4], ql[4][4];
    constexpr float kScale = 0.125f * 1.4426950408889634f;

    for (int kt = 0; kt < nk; kt += AF_BK) {
        const int nkt = min(AF_BK, nk - kt);
        const int k_rows = min(kr_cap, (nkt + 15) & ~15);
        if (kt > 0) __syncthreads();                                     // previous K/V tile fully consumed
        for (int idx = tid; idx < k_rows * 32; idx += nthr) {
            const int r = idx >> 5, c = idx & 7, which = (idx >> 3) & 3;    // which: 0 Kh, 1 Kl, 2 Vh, 3 Vl
            const size_t row = (size_t)(k0 + kt + min(r, nkt - 1)) * in.ldk + head * 64 + c * 8;
            const uint16_t* src = (which < 2) ? in.Kh + (which & 1 ? in.k_split : 0) + row : in.Vh + (which & 1 ? in.v_split : 0) + row;
            const uint32_t dst = (which == 0 ? sKh : which == 1 ? sKl : which == 2 ? sVh : sVl) + (uint32_t)(r * AF_S + c * 8) * 2u;
            cp_async16(dst, src, r < nkt);
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        __syncthreads();
        if (!warp_active) continue;
        if (kt == 0) {                                                   // A fragments of this warp's 16 query rows, 4 k-steps
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) {
                const uint32_t off = (uint32_t)((w * 16 + lm_row) * AF_S + 16 * ks + lm_col) * 2u;
                ldsm_x4(sQh + off, qh[ks][0], qh[ks][1], qh[ks][2], qh[ks][3]);
                ldsm_x4(sQl + off, ql[ks][0], ql[ks][1], ql[ks][2], ql[ks][3]);
            }
        }
        const int jmax = (nkt + 7) >> 3, smax = (nkt + 15) >> 4;

        // ---- S' = Q K^T * log2(e)/8.  One accumulator per 8-key block; per k-step the three passes (lo*hi, hi*lo, hi*hi) are issued
        // pass-major across all key blocks, so consecutive MMAs never depend on each other (8 independent chains). ----
        float s[8][4];
#pragma unroll
        for (int j = 0; j < 8; ++j) { s[j][0] = s[j][1] = s[j][2] = s[j][3] = 0.f; }
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {
#pragma unroll
            for (int jh = 0; jh < 2; ++jh) {                             // key-block pairs (2jh, 2jh+1): 4 independent chains per pass
                if (4 * jh < jmax) {
                    uint32_t kh[2][4], kl[2][4];                         // (b0,b1) of block 2jp, (b0,b1) of block 2jp+1
#pragma unroll
                    for (int q2 = 0; q2 < 2; ++q2) {
                        const int jp = 2 * jh + q2;
                        const uint32_t off = (uint32_t)((16 * jp + lk_row) * AF_S + 16 * ks + lk_col) * 2u;
                        ldsm_x4(sKh + off, kh[q2][0], kh[q2][1], kh[q2][2], kh[q2][3]);
                        ldsm_x4(sKl + off, kl[q2][0], kl[q2][1], kl[q2][2], kl[q2][3]);
                    }
#pragma unroll
                    for (int q2 = 0; q2 < 2; ++q2) { const int jp = 2 * jh + q2; mma_f16(s[2 * jp], ql[ks], kh[q2][0], kh[q2][1]); mma_f16(s[2 * jp + 1], ql[ks], kh[q2][2], kh[q2][3]); }
#pragma unroll
                    for (int q2 = 0; q2 < 2; ++q2) { const int jp = 2 * jh + q2; mma_f16(s[2 * jp], qh[ks], kl[q2][0], kl[q2][1]); mma_f16(s[2 * jp + 1], qh[ks], kl[q2][2], kl[q2][3]); }
#pragma unroll
                    for (int q2 = 0; q2 < 2; ++q2) { const int jp = 2 * jh + q2; mma_f16(s[2 * jp], qh[ks], kh[q2][0], kh[q2][1]); mma_f16(s[2 * jp + 1], qh[ks], kh[q2][2], kh[q2][3]); }
                }
            }
        }
#pragma unroll
        for (int jp = 0; jp < 4; ++jp) {
            if (2 * jp < jmax) {
#pragma unroll
                for (int i = 0; i < 4; ++i) { s[2 * jp][i] *= kScale; s[2 * jp + 1][i] *= kScale; }
                if (nkt < AF_BK) {                                       // only the last, partial tile has keys to mask
                    const int key = 16 * jp + 2 * t;
                    if (key >= nkt) { s[2 * jp][0] = -INFINITY; s[2 * jp][2] = -INFINITY; }
                    if (key + 1 >= nkt) { s[2 * jp][1] = -INFINITY; s[2 * jp][3] = -INFINITY; }
                    if (key + 8 >= nkt) { s[2 * jp + 1][0] = -INFINITY; s[2 * jp + 1][2] = -INFINITY; }
                    if (key + 9 >= nkt) { s[2 * jp + 1][1] = -INFINITY; s[2 * jp + 1][3] = -INFINITY; }
                }
            } else {
#pragma unroll
                for (int i = 0; i < 4; ++i) { s[2 * jp][i] = -INFINITY; s[2 * jp + 1][i] = -INFINITY; }
            }
        }
        // ---- online softmax ----
        float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
        for (int j = 0; j < 8; ++j) { mx0 = fmaxf(mx0, fmaxf(s[j][0], s[j][1])); mx1 = fmaxf(mx1, fmaxf(s[j][2], s[j][3])); }
        mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1)); mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2));
        mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1)); mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
        const float mn0 = fmaxf(m0, mx0), mn1 = fmaxf(m1, mx1);          // finite: every tile has >= 1 valid key
        const float corr0 = ex2_approx(m0 - mn0), corr1 = ex2_approx(m1 - mn1);
        const float b0 = 12.0f - mn0, b1 = 12.0f - mn1;
        float ps0 = 0.f, ps1 = 0.f;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            s[j][0] = ex2_approx(s[j][0] + b0); s[j][1] = ex2_approx(s[j][1] + b0);
            s[j][2] = ex2_approx(s[j][2] + b1); s[j][3] = ex2_approx(s[j][3] + b1);
            ps0 += s[j][0] + s[j][1]; ps1 += s[j][2] + s[j][3];
        }
        ps0 += __shfl_xor_sync(0xffffffffu, ps0, 1); ps0 += __shfl_xor_sync(0xffffffffu, ps0, 2);
        ps1 += __shfl_xor_sync(0xffffffffu, ps1, 1); ps1 += __shfl_xor_sync(0xffffffffu, ps1, 2);
        l0s = l0s * corr0 + ps0; l1s = l1s * corr1 + ps1;
        m0 = mn0; m1 = mn1;
#pragma unroll
        for (int u = 0; u < 8; ++u) { o[u][0] *= corr0; o[u][1] *= corr0; o[u][2] *= corr1; o[u][3] *= corr1; }

        // ---- O' += P' V (accumulated straight into the rescaled running output), pass-major across the 8 output tiles ----
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {
            if (ks < smax) {
                uint32_t pa[4], pl[4];                                   // A fragment = S accumulators of key blocks 2ks, 2ks+1
                split_pair(s[2 * ks][0], s[2 * ks][1], pa[0], pl[0]);
                split_pair(s[2 * ks][2], s[2 * ks][3], pa[1], pl[1]);
                split_pair(s[2 * ks + 1][0], s[2 * ks + 1][1], pa[2], pl[2]);
                split_pair(s[2 * ks + 1][2], s[2 * ks + 1][3], pa[3], pl[3]);
                const uint32_t voff = (uint32_t)((16 * ks + lm_row) * AF_S + lm_col) * 2u;
#pragma unroll
                for (int uh = 0; uh < 2; ++uh) {                         // output tiles 4uh .. 4uh+3: 4 independent chains per pass
                    uint32_t vh[2][4], vl[2][4];
#pragma unroll
                    for (int q2 = 0; q2 < 2; ++q2) {
                        ldsm_x4_t(sVh + voff + (2 * uh + q2) * 32, vh[q2][0], vh[q2][1], vh[q2][2], vh[q2][3]);
                        ldsm_x4_t(sVl + voff + (2 * uh + q2) * 32, vl[q2][0], vl[q2][1], vl[q2][2], vl[q2][3]);
                    }
#pragma unroll
                    for (int q2 = 0; q2 < 2; ++q2) { const int up = 2 * uh + q2; mma_f16(o[2 * up], pl, vh[q2][0], vh[q2][1]); mma_f16(o[2 * up + 1], pl, vh[q2][2], vh[q2][3]); }
#pragma unroll
                    for (int q2 = 0; q2 < 2; ++q2) { const int up = 2 * uh + q2; mma_f16(o[2 * up], pa, vl[q2][0], vl[q2][1]); mma_f16(o[2 * up + 1], pa, vl[q2][2], vl[q2][3]); }
#pragma unroll
                    for (int q2 = 0; q2 < 2; ++q2) { const int up = 2 * uh + q2; mma_f16(o[2 * up], pa, vh[q2][0], vh[q2][1]); mma_f16(o[2 * up + 1], pa, vh[q2][2], vh[q2][3]); }
                }
            }
        }
    }
    if (!warp_active) return;

    const float inv0 = 1.0f / l0s, inv1 = 1.0f / l1s;
    const bool odd = (t & 1) != 0;
#pragma unroll
    for (int half = 0; half < 2; ++half) {
        const int r = w * 16 + gq + half * 8;
        const float inv = half ? inv1 : inv0;
        const int e = half * 2;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            const float a0 = o[2 * c][e] * inv, a1 = o[2 * c][e + 1] * inv, b0 = o[2 * c + 1][e] * inv, b1 = o[2 * c + 1][e + 1] * inv;
            const float s0 = odd ? a0 : b0, s1 = odd ? a1 : b1;
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

// ------------------------------------------------------------------------------------------------------------------------
// attn_last_kernel: self-attention of the LAST prefix position only (the pruned last decoder layer, FFB_OPT_PRUNE_LAST: only
// pointer[-1] is consumed, model_para.py:176).  One CTA per sequence, one warp per head: the query row b*P + P-1 against the P
// keys / values of the sequence, all read as fp16x2 halves from the q|k|v operand buffer and recombined to fp32 (hi + lo), plain
// fp32 dot products (lane = 2 head dims, warp-shuffle reduction), online softmax, output row b as fp16x2.
// (The general kernels spend a whole CTA - staging, fragments, MMAs - on this one query row: 187 us per launch vs ~40 us here.)
// ------------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) attn_last_kernel(const uint16_t* __restrict__ qkv, long long split_stride, int ld, int E,
                                                        int P, int B, uint16_t* __restrict__ Os, long long os_stride, int ldo,
                                                        const int* stop, const uint16_t* __restrict__ qlast, long long ql_split) {
    FFB_PDL_SYNC();
    FFB_STOP_CHECK(stop);
    const int b = blockIdx.x, head = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (b >= B || head * 64 >= E) return;
    auto ld2 = [&](const uint16_t* base, long long row, int col) -> float2 {       // (hi + lo) of two adjacent elements
        const __half2 h = *reinterpret_cast<const __half2*>(base + row * ld + col);
        const __half2 l = *reinterpret_cast<const __half2*>(base + split_stride + row * ld + col);
        const float2 hf = __half22float2(h), lf = __half22float2(l);
        return make_float2(hf.x + lf.x, hf.y + lf.y);
    };
    const long long r0 = (long long)b * P;
    const int c = head * 64 + 2 * lane;
    // q: row b of the compact last-position buffer [2][B][E] (the q projection of the last layer is computed for these rows only),
    // or the q columns of the full q|k|v buffer
    float2 q;
    if (qlast != nullptr) {
        const float2 hf = __half22float2(*reinterpret_cast<const __half2*>(qlast + (size_t)b * E + c));
        const float2 lf = __half22float2(*reinterpret_cast<const __half2*>(qlast + ql_split + (size_t)b * E + c));
        q = make_float2(hf.x + lf.x, hf.y + lf.y);
    } else q = ld2(qkv, r0 + P - 1, c);
    constexpr float kScale = 0.125f * 1.4426950408889634f;                        // scores in the log2 domain
    float m = -INFINITY, l = 0.f, o0 = 0.f, o1 = 0.f;
    for (int j = 0; j < P; ++j) {
        const float2 k = ld2(qkv, r0 + j, E + c);
        const float2 v = ld2(qkv, r0 + j, 2 * E + c);
        const float s = warp_sum(q.x * k.x + q.y * k.y) * kScale;
        const float mn = fmaxf(m, s);
        const float corr = ex2_approx(m - mn), p = ex2_approx(s - mn);
        l = l * corr + p;
        o0 = o0 * corr + p * v.x; o1 = o1 * corr + p * v.y;
        m = mn;
    }
    const float inv = 1.0f / l;
    uint32_t hw, lw;
    split_pair(o0 * inv, o1 * inv, hw, lw);
    *reinterpret_cast<uint32_t*>(Os + (size_t)b * ldo + c) = hw;
    *reinterpret_cast<uint32_t*>(Os + os_stride + (size_t)b * ldo + c) = lw;
}

}  // namespace ffb
