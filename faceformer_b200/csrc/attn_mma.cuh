// attn_mma.cuh -- attention core on the tensor pipe with fp32-class accuracy (3xTF32 via mma.sync m16n8k8).
//
//   O = softmax(Q K^T / 8) V   per (group, head), head dim 64, any number of keys (online softmax, 64-key tiles)
//
// Replaces both SIMT attention kernels (kernels.cuh: attn_rows_kernel for the decoder self-attention,
// attn_tiled_kernel for encoder self-attention / decoder cross-attention; reference call sites
// transformer.py:169-171,244-251 through torch F.multi_head_attention_forward).
//
// Precision: every product is done as hi*hi + (lo*hi + hi*lo) with hi = the operand as the tensor core reads it
// (fp32 truncated to TF32) and lo = RN_tf32(x - hi): relative error ~2^-21, below fp32 accumulation noise.
// The dominant hi*hi products and the small corrections use separate register accumulators, and every 64-key
// tile is accumulated from zero and added to the running output in fp32 (round-to-nearest), so truncating
// tensor-core adds never form long chains.
//
// Layout tricks (all exact re-indexings of the contraction / output dimensions):
//   * the 64 head dims are assigned to the MMA k-slots so that a thread's A/B fragment elements are 16
//     contiguous floats in shared memory (4 x LDS.128 instead of 32 x LDS.32), conflict-free with row stride 80;
//   * the S accumulator layout of key tile j is reused AS IS as the A fragment of the P.V product
//     (k-slot t <-> key 8j+2t, slot t+4 <-> key 8j+2t+1), so P never leaves registers;
//   * output columns of the P.V product are permuted so that each thread loads V as LDS.128 (row stride 68,
//     conflict-free) and finally owns two runs of 8 contiguous output floats.
#pragma once
#include "kernels.cuh"

namespace ffb {

constexpr int AM_BQ = 64, AM_BK = 64, AM_SQ = 80, AM_SV = 68;
constexpr int AM_SMEM_BYTES = (2 * AM_BQ * AM_SQ + AM_BK * AM_SV) * (int)sizeof(float);   // Q, K (stride 80), V (stride 68)

__device__ __forceinline__ void mma_tf32(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile(
        "mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
        : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
// lo part of the hi/lo TF32 split; the hi part is the raw fp32 word (the tensor core ignores its low 13 bits)
__device__ __forceinline__ uint32_t tf32_lo(float x) {
    const float hi = __uint_as_float(__float_as_uint(x) & 0xffffe000u);
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x - hi));
    return r;
}

// One work item = (group, head, tile of <= 64 query rows) on 128 threads; `smem` = AM_SMEM_BYTES of shared memory.  Written as a device routine
// so that a multi-phase kernel can run it on any 128-thread group: global reads go through L2 (__ldcg: q / k / v may have been written by other
// CTAs of the SAME launch), the 128 threads meet at named barrier `bar_id`, and with q_staged the caller has already placed the pre-scaled Q tile in
// smem[0, 64 * 80).  (persist.cuh started from this routine and now carries its own warp-level variant of the same arithmetic, attn_item.)
__device__ __forceinline__ void attn_mma_tile(float* smem, bool q_staged, const float* __restrict__ Q, int ldq,
                                              const float* __restrict__ K, const float* __restrict__ V, int ldk,
                                              float* __restrict__ O, int ldo, uint16_t* __restrict__ Os, long long os_stride,
                                              int split_fmt, int* overflow, long long q0, int nqt, long long k0, int nk, long long o0, int head,
                                              int tid, int bar_id) {
    // the 128 threads of this work item meet at named barrier `bar_id` (0 with a 128-thread CTA = __syncthreads)
    auto sync128 = [&]() { asm volatile("bar.sync %0, 128;" ::"r"(bar_id) : "memory"); };
    float* Qs = smem;                              // [64][80]
    float* Ks = smem + AM_BQ * AM_SQ;              // [64][80]
    float* Vs = smem + 2 * AM_BQ * AM_SQ;          // [64][68]
    const int lane = tid & 31, w = tid >> 5;
    const int gq = lane >> 2, t = lane & 3;        // MMA fragment coordinates

    if (!q_staged) {
        sync128();                                 // a previous work item of this CTA may still be reading the tile buffers
        // stage the Q tile, pre-scaled by sqrt(1/64) = 0.125 (exact)
        for (int idx = tid; idx < AM_BQ * 16; idx += 128) {
            const int r = idx >> 4, d4 = idx & 15;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (r < nqt) {
                v = __ldcg(reinterpret_cast<const float4*>(Q + (size_t)(q0 + r) * ldq + head * 64 + d4 * 4));
                v.x *= 0.125f; v.y *= 0.125f; v.z *= 0.125f; v.w *= 0.125f;
            }
            *reinterpret_cast<float4*>(Qs + r * AM_SQ + d4 * 4) = v;
        }
    }
    sync128();

    const bool warp_active = (w * 16) < nqt;       // warps whose 16 rows are all padding skip the math
    // A fragments of this warp's 16 query rows for all 8 k-steps: chunk c (LDS.128) holds k-steps 2c, 2c+1
    uint32_t qa[8][4], ql[8][4];
    if (warp_active) {
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            const float4 r0 = *reinterpret_cast<const float4*>(Qs + (w * 16 + gq) * AM_SQ + 16 * c + 4 * t);
            const float4 r1 = *reinterpret_cast<const float4*>(Qs + (w * 16 + gq + 8) * AM_SQ + 16 * c + 4 * t);
            // k-step 2c: slots (t, t+4) = elements (x, y); k-step 2c+1: (z, w).  a0=(g,t) a1=(g+8,t) a2=(g,t+4) a3=(g+8,t+4)
            qa[2 * c][0] = __float_as_uint(r0.x); qa[2 * c][1] = __float_as_uint(r1.x);
            qa[2 * c][2] = __float_as_uint(r0.y); qa[2 * c][3] = __float_as_uint(r1.y);
            qa[2 * c + 1][0] = __float_as_uint(r0.z); qa[2 * c + 1][1] = __float_as_uint(r1.z);
            qa[2 * c + 1][2] = __float_as_uint(r0.w); qa[2 * c + 1][3] = __float_as_uint(r1.w);
            ql[2 * c][0] = tf32_lo(r0.x); ql[2 * c][1] = tf32_lo(r1.x); ql[2 * c][2] = tf32_lo(r0.y); ql[2 * c][3] = tf32_lo(r1.y);
            ql[2 * c + 1][0] = tf32_lo(r0.z); ql[2 * c + 1][1] = tf32_lo(r1.z);
            ql[2 * c + 1][2] = tf32_lo(r0.w); ql[2 * c + 1][3] = tf32_lo(r1.w);
        }
    }

    float m0 = -INFINITY, m1 = -INFINITY, l0 = 0.f, l1 = 0.f;     // rows gq and gq+8 of this warp
    float o[8][4];                                                // o[u]: rows (gq, gq+8) x output cols (2t, 2t+1) of tile u
#pragma unroll
    for (int u = 0; u < 8; ++u) { o[u][0] = o[u][1] = o[u][2] = o[u][3] = 0.f; }

    for (int kt = 0; kt < nk; kt += AM_BK) {
        const int nkb = min(8, (nk - kt + 7) >> 3);               // 8-key blocks of this tile that hold keys: the others are skipped
        sync128();                                                // previous K/V tile fully consumed
        for (int idx = tid; idx < AM_BK * 16; idx += 128) {
            const int r = idx >> 4, d4 = idx & 15;
            float4 kv = make_float4(0.f, 0.f, 0.f, 0.f), vv = kv;
            if (kt + r < nk) {
                const size_t off = (size_t)(k0 + kt + r) * ldk + head * 64 + d4 * 4;
                kv = __ldcg(reinterpret_cast<const float4*>(K + off));
                vv = __ldcg(reinterpret_cast<const float4*>(V + off));
            }
            *reinterpret_cast<float4*>(Ks + r * AM_SQ + d4 * 4) = kv;
            *reinterpret_cast<float4*>(Vs + r * AM_SV + d4 * 4) = vv;
        }
        sync128();
        if (!warp_active) continue;

        // ---- S = Q K^T for this warp's 16 rows x 64 keys: s[j] = key tile j (keys 8j..8j+7), C layout ----
        float s[8][4];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            float sm[4] = {0.f, 0.f, 0.f, 0.f}, sc[4] = {0.f, 0.f, 0.f, 0.f};
            if (j < nkb) {
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                const float4 kb = *reinterpret_cast<const float4*>(Ks + (8 * j + gq) * AM_SQ + 16 * c + 4 * t);
                // k-step 2c: b0 = slot t -> kb.x, b1 = slot t+4 -> kb.y ; k-step 2c+1: kb.z, kb.w
                mma_tf32(sc, ql[2 * c], __float_as_uint(kb.x), __float_as_uint(kb.y));
                mma_tf32(sc, qa[2 * c], tf32_lo(kb.x), tf32_lo(kb.y));
                mma_tf32(sm, qa[2 * c], __float_as_uint(kb.x), __float_as_uint(kb.y));
                mma_tf32(sc, ql[2 * c + 1], __float_as_uint(kb.z), __float_as_uint(kb.w));
                mma_tf32(sc, qa[2 * c + 1], tf32_lo(kb.z), tf32_lo(kb.w));
                mma_tf32(sm, qa[2 * c + 1], __float_as_uint(kb.z), __float_as_uint(kb.w));
            }
            }
            // C layout: [0]=(gq, 2t) [1]=(gq, 2t+1) [2]=(gq+8, 2t) [3]=(gq+8, 2t+1); mask keys beyond nk
            const int key = kt + 8 * j + 2 * t;
            s[j][0] = (key < nk) ? sm[0] + sc[0] : -INFINITY;
            s[j][1] = (key + 1 < nk) ? sm[1] + sc[1] : -INFINITY;
            s[j][2] = (key < nk) ? sm[2] + sc[2] : -INFINITY;
            s[j][3] = (key + 1 < nk) ? sm[3] + sc[3] : -INFINITY;
        }
        // ---- online softmax (rows live in the 4 lanes of a quad) ----
        float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
        for (int j = 0; j < 8; ++j) { mx0 = fmaxf(mx0, fmaxf(s[j][0], s[j][1])); mx1 = fmaxf(mx1, fmaxf(s[j][2], s[j][3])); }
        mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1)); mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2));
        mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1)); mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
        const float mn0 = fmaxf(m0, mx0), mn1 = fmaxf(m1, mx1);   // finite: every tile has >= 1 valid key
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
        l0 = l0 * corr0 + ps0; l1 = l1 * corr1 + ps1;
        m0 = mn0; m1 = mn1;

        // ---- O_tile = P V, accumulated from zero, then O = O * corr + O_tile (fp32, round to nearest) ----
        // k-step j contracts over keys 8j..8j+7 with slot t <-> key 8j+2t, slot t+4 <-> key 8j+2t+1, so the A fragment
        // is the S accumulator itself: a0 = s[j][0], a1 = s[j][2], a2 = s[j][1], a3 = s[j][3].
        // Output tile u, fragment column n <-> head dim 4n + (u & 3) + 32 (u >> 2): thread (n = gq) reads V[key][4 gq .. 4 gq + 3]
        // (u = 0..3) and V[key][32 + 4 gq ..] (u = 4..7) as two LDS.128 per key row.
        float om[8][4], oc[8][4];
#pragma unroll
        for (int u = 0; u < 8; ++u) { om[u][0] = om[u][1] = om[u][2] = om[u][3] = 0.f; oc[u][0] = oc[u][1] = oc[u][2] = oc[u][3] = 0.f; }
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            if (j >= nkb) continue;                               // all probabilities of the block are exactly 0
            uint32_t pa[4], pl[4];
            pa[0] = __float_as_uint(s[j][0]); pa[1] = __float_as_uint(s[j][2]); pa[2] = __float_as_uint(s[j][1]); pa[3] = __float_as_uint(s[j][3]);
            pl[0] = tf32_lo(s[j][0]); pl[1] = tf32_lo(s[j][2]); pl[2] = tf32_lo(s[j][1]); pl[3] = tf32_lo(s[j][3]);
            const float* v0 = Vs + (8 * j + 2 * t) * AM_SV + 4 * gq;          // key of slot t
            const float* v1 = v0 + AM_SV;                                      // key of slot t+4
            const float4 a_lo = *reinterpret_cast<const float4*>(v0), a_hi = *reinterpret_cast<const float4*>(v0 + 32);
            const float4 b_lo = *reinterpret_cast<const float4*>(v1), b_hi = *reinterpret_cast<const float4*>(v1 + 32);
            const float e0[8] = {a_lo.x, a_lo.y, a_lo.z, a_lo.w, a_hi.x, a_hi.y, a_hi.z, a_hi.w};
            const float e1[8] = {b_lo.x, b_lo.y, b_lo.z, b_lo.w, b_hi.x, b_hi.y, b_hi.z, b_hi.w};
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                mma_tf32(oc[u], pl, __float_as_uint(e0[u]), __float_as_uint(e1[u]));
                mma_tf32(oc[u], pa, tf32_lo(e0[u]), tf32_lo(e1[u]));
                mma_tf32(om[u], pa, __float_as_uint(e0[u]), __float_as_uint(e1[u]));
            }
        }
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            o[u][0] = o[u][0] * corr0 + (om[u][0] + oc[u][0]); o[u][1] = o[u][1] * corr0 + (om[u][1] + oc[u][1]);
            o[u][2] = o[u][2] * corr1 + (om[u][2] + oc[u][2]); o[u][3] = o[u][3] * corr1 + (om[u][3] + oc[u][3]);
        }
    }
    if (!warp_active) return;

    // thread owns, for row gq (o[u][0..1]) and row gq+8 (o[u][2..3]): head dims 8t + {0..3} (u=0..3, col 2t), 8t + 4 + {0..3} (col 2t+1),
    // and the same +32 for u = 4..7  ->  two runs of 8 contiguous floats per row.
    const float inv0 = 1.0f / l0, inv1 = 1.0f / l1;
#pragma unroll
    for (int half = 0; half < 2; ++half) {
        const int r = w * 16 + gq + half * 8;
        if (r >= nqt) continue;
        const float inv = half ? inv1 : inv0;
        const int e = half * 2;
#pragma unroll
        for (int hi = 0; hi < 2; ++hi) {                           // u = 4*hi .. 4*hi+3  -> head dims 32*hi + 8t + ...
            const float4 c0 = make_float4(o[4 * hi][e] * inv, o[4 * hi + 1][e] * inv, o[4 * hi + 2][e] * inv, o[4 * hi + 3][e] * inv);
            const float4 c1 = make_float4(o[4 * hi][e + 1] * inv, o[4 * hi + 1][e + 1] * inv, o[4 * hi + 2][e + 1] * inv, o[4 * hi + 3][e + 1] * inv);
            const size_t off = (size_t)(o0 + r) * ldo + head * 64 + 32 * hi + 8 * t;
            if (Os == nullptr) {
                *reinterpret_cast<float4*>(O + off) = c0;
                *reinterpret_cast<float4*>(O + off + 4) = c1;
            } else {
                store_split4(Os + off, os_stride, c0, split_fmt, overflow);
                store_split4(Os + off + 4, os_stride, c1, split_fmt, overflow);
            }
        }
    }
}

__global__ void __launch_bounds__(128, 2) attn_mma_kernel(const float* __restrict__ Q, int ldq,
                                                          const float* __restrict__ K, const float* __restrict__ V, int ldk,
                                                          float* __restrict__ O, int ldo, uint16_t* __restrict__ Os,
                                                          long long os_stride, const AttnGroups g, const int* stop) {
    FFB_STOP_CHECK(stop);
    extern __shared__ __align__(16) float smem[];
    long long q0, k0, o0; int nq, nk;
    attn_group(g, blockIdx.x, q0, nq, k0, nk, o0);
    const int qt0 = blockIdx.z * AM_BQ;
    if (qt0 >= nq) return;
    attn_mma_tile(smem, false, Q, ldq, K, V, ldk, O, ldo, Os, os_stride, g.split_fmt, g.overflow, q0 + qt0, min(AM_BQ, nq - qt0), k0, nk, o0 + qt0,
                  (int)blockIdx.y, (int)threadIdx.x, 0);
}

}  // namespace ffb
