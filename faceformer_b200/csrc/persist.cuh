// persist.cuh -- the whole greedy decode loop of a SMALL batch as ONE persistent cooperative kernel (sm_100a).
//
// Reference: SurfaceFormer_Parallel.forward_eval (faceformer/models/model_para.py:216-236) and SurfaceFormer.forward_eval
// (faceformer/models/model.py:192-213) over TransformerDecoderLayer.forward_pre (faceformer/transformer.py:235-256), decoder.norm
// (transformer.py:115-116), project + select_next (model_para.py:173-179,225).
//
// Why: with one wireframe per batch (the reference's own test loop, trainer.py:51; BASELINE configs[0]) a decode step is 65 MB of weights
// against a few hundred activation rows.  As ~72 dependent kernel launches it costs 0.7-0.9 ms; the work itself is a few microseconds per
// phase.  This kernel runs ALL steps of the loop in one launch: the grid (one CTA of 256 threads per SM, co-resident through a cooperative
// launch) walks the phases of a step and meets at a grid-wide barrier between them; the stop predicate is evaluated on the device after
// every step.
//
// Phases per decoder layer (10) + the head (1):
//   0  xs, xps = split(LN1(x)), split(LN1(x) + qpos)            layer 0: x = memory[tokens] is gathered here (model_para.py:217-219)
//   1  q,k,v   = [xps | xps | xs] W_in^T + b
//   2  att     = softmax(q k^T / 8) v  per (sequence, head)     no causal mask (model_para.py:222-223); written as fp16x2
//   3  x      += att W_out^T + b
//   4  xps     = split(LN2(x) + qpos)
//   5  att     = softmax(q_c Kc^T / 8) Vc  with  q_c = xps W_q^T + b     one item = (32 rows, one head): the q tile never leaves the CTA
//   6  x      += att W_out^T + b
//   7  xs      = split(LN3(x))
//   8  h       = split(relu(xs W_1^T + b))
//   9  x      += h W_2^T + b
//   head       float64 LayerNorm of the last position, folded pointer dot (memW, ffb200.cu run_head_fold), first-max argmax, append, counters
//
// GEMM items are 32 x 32 output tiles (32 x 64 for the fused phase 5) on 256 threads: two groups of four warps split K (even / odd 64-wide
// chunks) and are summed in a fixed order; mma.sync m16n8k16 on fp16x2 split operands (hi*hi + lo*hi + hi*lo), fp32 accumulation, register
// accumulators drained with round-to-nearest adds every 128 k of a group -- the precision class of gemm_tc.cuh.  Every GEMM operand lives in
// global memory (L2 at these sizes) already split, so a chunk goes cp.async -> ldmatrix -> MMA with no conversion; up to 192 KB (a whole
// K = 512 item) is in flight per CTA; bias / residual are prefetched before the mainloop.  Weights are the pre-scaled fp16x2 arrays of the
// tensor-core path.  Attention items use all 8 warps (attn_item below): the arithmetic of attn_mma.cuh (3xTF32 mma.sync, online softmax), keys
// split over the warps, partials merged in a fixed order.
//
// Performance notes (profiles/persist_phase_clocks_r2.md).  One L2 access costs ~1.1 k clk in situ and the per-barrier acquire fence drops
// L1, so every phase is a chain of L2 round trips and every spilled register is one more: the kernel is built to stay spill-free in its hot
// loops at 255 registers (one call site per routine, arrays sized by the NV template parameter).  It is sensitive to that: two later
// experiments that added a third GEMM instantiation or a LayerNorm tail to the GEMM items lost 5-30 % through re-allocation alone.
#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "attn_h.cuh"     // cp_async16, ldsm_x4, mma_f16, split_pair
#include "attn_mma.cuh"   // attn_mma_tile
#include "enc64.cuh"      // warp_sum_d
#include "kernels.cuh"

namespace ffb {
namespace pd {

constexpr int TM = 32, TN = 64, TK = 64, NPAIR = 4, THREADS = 256;
constexpr int MAX_WF = 64, MAX_LAYERS = 8;
constexpr int SPLIT_A = TM * TK * 2;                       //  4 KB: fp16 rows of 128 B, 16-byte piece c of row r at c ^ (r & 7)
constexpr int SPLIT_W = TN * TK * 2;                       //  8 KB
constexpr int CHUNK_A = 2 * SPLIT_A, CHUNK_BYTES = CHUNK_A + 2 * SPLIT_W;     // 24 KB: A hi, A lo, W hi, W lo of one 64-wide k chunk
constexpr int PAIR_BYTES = 2 * CHUNK_BYTES;                // 48 KB: chunk 2i for warp group 0, chunk 2i + 1 for warp group 1
constexpr int OFF_RED = NPAIR * PAIR_BYTES;                // 192 KB ring, then the split-K hand-over buffer: float [128][16]
constexpr int OFF_TILE = OFF_RED + 128 * 16 * 4;           // int[MAX_WF + 1]: first tile of every wireframe
constexpr int OFF_WFI = OFF_TILE + (MAX_WF + 1) * 4;       // int[3][MAX_WF + 1]: seq_off, row_off, vlen (every global load costs an L2 round trip)
constexpr int OFF_MISC = OFF_WFI + 3 * (MAX_WF + 1) * 4;   // head phase: best value / index per warp
constexpr int SMEM_BYTES = ((OFF_MISC + 128 + 127) / 128) * 128;
static_assert(SMEM_BYTES <= 227 * 1024, "one CTA per SM");

struct LayerP {
    const uint16_t *w_sa_in, *w_sa_out, *w_ca_q, *w_ca_out, *w_l1, *w_l2;      // fp16x2 [2][N][K], pre-scaled by a power of two
    float s_sa_in, s_sa_out, s_ca_q, s_ca_out, s_l1, s_l2;                    // 1 / that power of two
    const float *b_sa_in, *b_sa_out, *b_ca_q, *b_ca_out, *b_l1, *b_l2;
    const float *n1w, *n1b, *n2w, *n2b, *n3w, *n3b;
};
struct Params {
    LayerP L[MAX_LAYERS];
    int Ld, E, FF, H, B, T, N, Lrows, mode, num_token;
    int max_steps;                                 // decode steps this launch may run (<= T - 1): the host continues with the per-step kernels
                                                   // behind it when the batch outgrows the latency-bound regime (rows = sequences x prefix length)
    float *x, *qkv;                                // fp32: residual stream [M, E]; q,k,v [M, 3E]; rows ordered (sequence, position)
    uint16_t *xs, *xps, *atts, *hs;                // fp16x2 GEMM operands [2][cap][E or FF]: LN(x), LN(x) + qpos, attention output, FFN hidden
    long long ssE, ssF;                            // elements between the two halves of a split array
    const float *mem, *memW, *Kc, *Vc, *qpos, *dec_nw, *dec_nb;
    const int *row_off, *vlen, *seq_wf, *seq_off;
    int* tok;                                      // [T][B] step-major tokens
    float* logits;                                 // [B][Lrows]
    int* state;                                    // ffb_handle::state
    unsigned* bar;                                 // grid barrier counter (zeroed before the launch)
    int* counts;                                   // [T] per-step counters (zeroed before the launch)
    long long* prof;                               // optional [2][16] clock sums of CTA 0 (work / barrier wait per phase kind; 10 = head)
};

struct GemmDesc {
    const uint16_t* A0; const uint16_t* A1; int n_switch;     // split A operand: columns >= n_switch read A1 (q,k from xps, v from xs)
    long long a_split; int lda; int K;
    const uint16_t* W; int N; float scale; const float* bias;
    int relu;
    float* C; int ldc;                             // fp32 output (or null)
    const float* R; int ldr;                       // residual added to the fp32 output (may alias C)
    uint16_t* Cs; long long cs_split; int ldcs;    // fp16x2 output (or null)
};

__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ uint32_t s_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// Grid-wide barrier: every CTA adds 1 (release) to a monotone counter and waits (acquire) until it reaches the CTA count times the number
// of barriers so far.  Data that crosses the barrier is always read through L2 (__ldcg / cp.async.cg), never through a stale L1 line.
__device__ __forceinline__ void grid_sync(unsigned* bar, unsigned& target) {
    __syncthreads();
    target += gridDim.x;
    if (threadIdx.x == 0) {
        asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(bar) : "memory");
        unsigned v;
        do { asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(bar) : "memory"); } while (v < target);
        asm volatile("fence.acq_rel.gpu;" ::: "memory");    // one acquire after the spin (an acquire load per poll invalidates L1 every iteration)
    }
    __syncthreads();
}

// tile t of the step -> (wireframe, first row, rows).  Tiles never straddle wireframes (the fused cross-attention needs one K / V set per tile).
__device__ __forceinline__ void tile_rows(const Params& p, const int* tile_first, int t, int P, int& wf, int& row0, int& nrows) {
    wf = 0;
    while (wf + 1 < p.N && tile_first[wf + 1] <= t) ++wf;
    const int* s_seq_off = tile_first + (MAX_WF + 1);
    const int first = s_seq_off[wf] * P, end = s_seq_off[wf + 1] * P;
    row0 = first + (t - tile_first[wf]) * TM;
    nrows = min(TM, end - row0);
}

// LayerNorm of every row (one warp per row; the arithmetic of layernorm_split_kernel) -> fp16x2 operands, optionally + qpos[row % P].
// gather (layer 0): the rows are the target embeddings memory[token] (model_para.py:217-219); they are also written to x.
template <int NV>     // float4 chunks per lane: E <= 128 * NV
__device__ __forceinline__ void ln_phase(const Params& p, int P, bool gather, const float* g, const float* b, uint16_t* out_plain, uint16_t* out_pos) {
    const int lane = threadIdx.x & 31, E = p.E, M = p.B * P;
    const int nv = E >> 7;
    const bool probe = p.prof && blockIdx.x == 0 && threadIdx.x == 0 && !gather;
    if (probe) {      // raw in-situ latencies: one L2 load of a freshly written activation, one read-only load of a parameter
        long long a0 = clock64();
        float v0 = __ldcg(p.x + 5 * p.E + 17);
        asm volatile("" ::"f"(v0) : "memory");
        long long a1 = clock64();
        float v1 = __ldg(g + 33);
        asm volatile("" ::"f"(v1) : "memory");
        long long a2 = clock64();
        float v2 = __ldcg(p.x + 7 * p.E + 19);
        asm volatile("" ::"f"(v2) : "memory");
        long long a3 = clock64();
        p.prof[24] += a1 - a0; p.prof[25] += a2 - a1; p.prof[23] += a3 - a2;
    }
    // row r -> CTA r % grid, warp r / grid: a few hundred rows spread over all SMs (one or two warps each) instead of filling 8 warps of a few
    for (int r = blockIdx.x + gridDim.x * (threadIdx.x >> 5); r < M; r += gridDim.x * (THREADS / 32)) {
        const long long c0 = probe ? clock64() : 0;
        const int pos = r % P;
        const float* xr;
        if (gather) { const int sq = r / P; xr = p.mem + (size_t)(p.row_off[p.seq_wf[sq]] + __ldcg(p.tok + (size_t)pos * p.B + sq)) * E; }
        else xr = p.x + (size_t)r * E;
        const float* prow = p.qpos + (size_t)pos * E;
        float4 v[NV], gm[NV], bt[NV], pp[NV];
#pragma unroll
        for (int i = 0; i < NV; ++i) {
            if (i < nv) {
                const int c = i * 128 + lane * 4;
                v[i] = __ldcg(reinterpret_cast<const float4*>(xr + c));
                gm[i] = __ldg(reinterpret_cast<const float4*>(g + c)); bt[i] = __ldg(reinterpret_cast<const float4*>(b + c));
                if (out_pos) pp[i] = __ldg(reinterpret_cast<const float4*>(prow + c));
            }
        }
        float s = 0.f;
#pragma unroll
        for (int i = 0; i < NV; ++i) {
            if (i < nv) {
                if (gather) *reinterpret_cast<float4*>(p.x + (size_t)r * E + i * 128 + lane * 4) = v[i];
                s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
            }
        }
        const long long c1 = probe ? clock64() : 0;
        const float mean = warp_sum(s) / (float)E;
        float q = 0.f;
#pragma unroll
        for (int i = 0; i < NV; ++i) {
            if (i < nv) {
                v[i].x -= mean; v[i].y -= mean; v[i].z -= mean; v[i].w -= mean;
                q += (v[i].x * v[i].x + v[i].y * v[i].y) + (v[i].z * v[i].z + v[i].w * v[i].w);
            }
        }
        const float rstd = 1.0f / sqrtf(warp_sum(q) / (float)E + 1e-5f);
        const long long c2 = probe ? clock64() : 0;
#pragma unroll
        for (int i = 0; i < NV; ++i) {
            if (i < nv) {
                const int c = i * 128 + lane * 4;
                float4 o;
                o.x = v[i].x * rstd * gm[i].x + bt[i].x; o.y = v[i].y * rstd * gm[i].y + bt[i].y;
                o.z = v[i].z * rstd * gm[i].z + bt[i].z; o.w = v[i].w * rstd * gm[i].w + bt[i].w;
                if (out_plain) store_split4(out_plain + (size_t)r * E + c, p.ssE, o, 2, p.state + 4);
                if (out_pos) {
                    o.x += pp[i].x; o.y += pp[i].y; o.z += pp[i].z; o.w += pp[i].w;
                    store_split4(out_pos + (size_t)r * E + c, p.ssE, o, 2, p.state + 4);
                }
            }
        }
        if (probe) { const long long c3 = clock64(); p.prof[28] += c1 - c0; p.prof[29] += c2 - c1; p.prof[30] += c3 - c2; p.prof[31] += 1; }
    }
}

// ---- attention work item on the whole CTA (8 warps) --------------------------------------------------------------------------------
// nq_sub (1, 2 or 4) sub-tiles of 16 query rows; the 8 / nq_sub warps of a sub-tile take the 32-key tiles of the key range round-robin
// (online softmax over their own tiles: the arithmetic of attn_mma.cuh, 3xTF32 mma.sync), then the partial (max, sum, O) triples of a
// sub-tile are merged in a fixed warp order.  A warp stages its own K / V tile: no CTA barrier inside the key loop, so long prefixes
// (seq2seq) spread their keys over the warps and short ones (parallel model) their query rows.
constexpr int AK = 32;                                             // keys per warp tile
constexpr int AQ_BYTES = TM * AM_SQ * 4;                           // staged Q tile of the fused cross-attention item
constexpr int AKV_WARP = AK * AM_SQ * 4 + AK * AM_SV * 4;          // K [32][80] + V [32][68] per warp
constexpr int OFF_AKV = AQ_BYTES;
constexpr int AMG_WARP = 16 * 64 * 4 + 128;                        // un-normalised O [16][64] + max[16] + sum[16] per warp
constexpr int OFF_AMG = OFF_AKV + 8 * AKV_WARP;
static_assert(OFF_AMG + 8 * AMG_WARP <= OFF_RED, "attention buffers alias the operand ring");

// Q: row 0 of the item at the head's first column (Q_SMEM: the pre-scaled tile the fused GEMM left in shared memory, else global fp32).
template <bool Q_SMEM>
__device__ __forceinline__ void attn_item(uint8_t* smem, const float* Q, int ldq, int nq, int nq_sub, const float* __restrict__ K,
                                          const float* __restrict__ V, int ldk, long long k0, int nk, int head, uint16_t* Os, long long os_stride,
                                          int ldo, long long o0, int* ovf, long long* prof) {
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5, gq = lane >> 2, t = lane & 3;
    const int wps = 8 / nq_sub, sub = w / wps, kt_first = w % wps;
    const bool probe = prof && blockIdx.x == 0 && tid == 0;
    const long long a0 = probe ? clock64() : 0;
    long long a1 = a0, a2 = a0, a3 = a0;
    float* Ks = reinterpret_cast<float*>(smem + OFF_AKV + w * AKV_WARP);     // [32][80]
    float* Vs = Ks + AK * AM_SQ;                                              // [32][68]
    float* mo = reinterpret_cast<float*>(smem + OFF_AMG + w * AMG_WARP);     // [16][64], then max[16], sum[16]
    const bool sub_active = sub * 16 < nq;

    float m0 = -INFINITY, m1 = -INFINITY, l0 = 0.f, l1 = 0.f;     // rows gq and gq + 8 of the sub-tile
    float o[8][4];
#pragma unroll
    for (int u = 0; u < 8; ++u) { o[u][0] = o[u][1] = o[u][2] = o[u][3] = 0.f; }

    // a warp stages its own 32-key K / V tile with cp.async (rows beyond the key range are zero-filled)
    auto load_tile = [&](int kt) {
        const uint32_t ks_s = s_u32(Ks), vs_s = s_u32(Vs);
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            const int idx = lane + 32 * i, r = idx >> 4, d4 = idx & 15;
            const bool ok = kt + r < nk;
            const size_t off = (size_t)(k0 + (ok ? kt + r : 0)) * ldk + head * 64 + d4 * 4;
            cp_async16(ks_s + (r * AM_SQ + d4 * 4) * 4, K + off, ok);
            cp_async16(vs_s + (r * AM_SV + d4 * 4) * 4, V + off, ok);
        }
        cp_async_commit();
    };
    if (sub_active && kt_first * AK < nk) {
        load_tile(kt_first * AK);
        // A fragments of the 16 query rows for all 8 k-steps (layout of attn_mma.cuh: chunk c holds k-steps 2c, 2c + 1)
        uint32_t qa[8][4];                                        // (the lo parts are re-derived where they are used: 32 registers less)
        {
            const int ra = sub * 16 + gq, rb = ra + 8;
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                float4 r0 = make_float4(0.f, 0.f, 0.f, 0.f), r1 = r0;
                if (Q_SMEM) {
                    r0 = *reinterpret_cast<const float4*>(Q + (size_t)ra * ldq + 16 * c + 4 * t);
                    r1 = *reinterpret_cast<const float4*>(Q + (size_t)rb * ldq + 16 * c + 4 * t);
                } else {
                    if (ra < nq) r0 = __ldcg(reinterpret_cast<const float4*>(Q + (size_t)ra * ldq + 16 * c + 4 * t));
                    if (rb < nq) r1 = __ldcg(reinterpret_cast<const float4*>(Q + (size_t)rb * ldq + 16 * c + 4 * t));
                    r0.x *= 0.125f; r0.y *= 0.125f; r0.z *= 0.125f; r0.w *= 0.125f;        // sqrt(1 / 64), exact
                    r1.x *= 0.125f; r1.y *= 0.125f; r1.z *= 0.125f; r1.w *= 0.125f;
                }
                qa[2 * c][0] = __float_as_uint(r0.x); qa[2 * c][1] = __float_as_uint(r1.x);
                qa[2 * c][2] = __float_as_uint(r0.y); qa[2 * c][3] = __float_as_uint(r1.y);
                qa[2 * c + 1][0] = __float_as_uint(r0.z); qa[2 * c + 1][1] = __float_as_uint(r1.z);
                qa[2 * c + 1][2] = __float_as_uint(r0.w); qa[2 * c + 1][3] = __float_as_uint(r1.w);
            }
        }
        for (int kt = kt_first * AK; kt < nk; kt += wps * AK) {
            const int nkb = min(4, (nk - kt + 7) >> 3);           // 8-key blocks of this tile that hold keys
            if (kt != kt_first * AK) { __syncwarp(); load_tile(kt); }     // (the first tile was requested before the Q fragments were loaded)
            cp_async_wait<0>();
            __syncwarp();
            if (probe && kt == 0) a1 = clock64();
            // ---- S = Q K^T: 16 rows x 32 keys.  Three independent accumulators per 8-key block (hi*hi, lo*hi, hi*lo: chains of 8 MMAs); a full
            // tile runs without per-block branches so that the four blocks interleave ----
            float s[4][4];
            auto s_block = [&](int j) {
                float sm[4] = {0.f, 0.f, 0.f, 0.f}, sa[4] = {0.f, 0.f, 0.f, 0.f}, sb[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    const float4 kb = *reinterpret_cast<const float4*>(Ks + (8 * j + gq) * AM_SQ + 16 * c + 4 * t);
                    uint32_t la[4], lb[4];
#pragma unroll
                    for (int e = 0; e < 4; ++e) { la[e] = tf32_lo(__uint_as_float(qa[2 * c][e])); lb[e] = tf32_lo(__uint_as_float(qa[2 * c + 1][e])); }
                    mma_tf32(sa, la, __float_as_uint(kb.x), __float_as_uint(kb.y));
                    mma_tf32(sb, qa[2 * c], tf32_lo(kb.x), tf32_lo(kb.y));
                    mma_tf32(sm, qa[2 * c], __float_as_uint(kb.x), __float_as_uint(kb.y));
                    mma_tf32(sa, lb, __float_as_uint(kb.z), __float_as_uint(kb.w));
                    mma_tf32(sb, qa[2 * c + 1], tf32_lo(kb.z), tf32_lo(kb.w));
                    mma_tf32(sm, qa[2 * c + 1], __float_as_uint(kb.z), __float_as_uint(kb.w));
                }
                const int key = kt + 8 * j + 2 * t;
                s[j][0] = (key < nk) ? sm[0] + (sa[0] + sb[0]) : -INFINITY;
                s[j][1] = (key + 1 < nk) ? sm[1] + (sa[1] + sb[1]) : -INFINITY;
                s[j][2] = (key < nk) ? sm[2] + (sa[2] + sb[2]) : -INFINITY;
                s[j][3] = (key + 1 < nk) ? sm[3] + (sa[3] + sb[3]) : -INFINITY;
            };
            if (nkb == 4) {
#pragma unroll
                for (int j = 0; j < 4; ++j) s_block(j);
            } else {
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    if (j < nkb) s_block(j);
                    else { s[j][0] = s[j][1] = s[j][2] = s[j][3] = -INFINITY; }
                }
            }
            // ---- online softmax (rows live in the 4 lanes of a quad) ----
            float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
            for (int j = 0; j < 4; ++j) { mx0 = fmaxf(mx0, fmaxf(s[j][0], s[j][1])); mx1 = fmaxf(mx1, fmaxf(s[j][2], s[j][3])); }
            mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1)); mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2));
            mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1)); mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
            const float mn0 = fmaxf(m0, mx0), mn1 = fmaxf(m1, mx1);   // finite: every tile has >= 1 valid key
            const float corr0 = expf(m0 - mn0), corr1 = expf(m1 - mn1);
            float ps0 = 0.f, ps1 = 0.f;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                s[j][0] = expf(s[j][0] - mn0); s[j][1] = expf(s[j][1] - mn0);
                s[j][2] = expf(s[j][2] - mn1); s[j][3] = expf(s[j][3] - mn1);
                ps0 += s[j][0] + s[j][1]; ps1 += s[j][2] + s[j][3];
            }
            ps0 += __shfl_xor_sync(0xffffffffu, ps0, 1); ps0 += __shfl_xor_sync(0xffffffffu, ps0, 2);
            ps1 += __shfl_xor_sync(0xffffffffu, ps1, 1); ps1 += __shfl_xor_sync(0xffffffffu, ps1, 2);
            l0 = l0 * corr0 + ps0; l1 = l1 * corr1 + ps1;
            m0 = mn0; m1 = mn1;
            if (probe && kt == 0) a2 = clock64();
            // ---- O_tile = P V from zero, then O = O * corr + O_tile (fp32, round to nearest); head dims [0, 32) and [32, 64) in turn ----
#pragma unroll
            for (int uh = 0; uh < 2; ++uh) {
                float om[4][4], oc[4][4];
#pragma unroll
                for (int u = 0; u < 4; ++u) { om[u][0] = om[u][1] = om[u][2] = om[u][3] = 0.f; oc[u][0] = oc[u][1] = oc[u][2] = oc[u][3] = 0.f; }
                auto pv_block = [&](int j) {
                    uint32_t pa[4], pl[4];
                    pa[0] = __float_as_uint(s[j][0]); pa[1] = __float_as_uint(s[j][2]); pa[2] = __float_as_uint(s[j][1]); pa[3] = __float_as_uint(s[j][3]);
                    pl[0] = tf32_lo(s[j][0]); pl[1] = tf32_lo(s[j][2]); pl[2] = tf32_lo(s[j][1]); pl[3] = tf32_lo(s[j][3]);
                    const float* v0 = Vs + (8 * j + 2 * t) * AM_SV + 4 * gq + 32 * uh;
                    const float4 ea = *reinterpret_cast<const float4*>(v0), eb = *reinterpret_cast<const float4*>(v0 + AM_SV);
                    const float e0[4] = {ea.x, ea.y, ea.z, ea.w};
                    const float e1[4] = {eb.x, eb.y, eb.z, eb.w};
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        mma_tf32(oc[u], pl, __float_as_uint(e0[u]), __float_as_uint(e1[u]));
                        mma_tf32(oc[u], pa, tf32_lo(e0[u]), tf32_lo(e1[u]));
                        mma_tf32(om[u], pa, __float_as_uint(e0[u]), __float_as_uint(e1[u]));
                    }
                };
                if (nkb == 4) {
#pragma unroll
                    for (int j = 0; j < 4; ++j) pv_block(j);
                } else {
#pragma unroll
                    for (int j = 0; j < 4; ++j) if (j < nkb) pv_block(j);     // the probabilities of the other blocks are exactly 0
                }
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int uu = 4 * uh + u;
                    o[uu][0] = o[uu][0] * corr0 + (om[u][0] + oc[u][0]); o[uu][1] = o[uu][1] * corr0 + (om[u][1] + oc[u][1]);
                    o[uu][2] = o[uu][2] * corr1 + (om[u][2] + oc[u][2]); o[uu][3] = o[uu][3] * corr1 + (om[u][3] + oc[u][3]);
                }
            }
        }
    }
    if (probe) a3 = clock64();
    // partial (max, sum, un-normalised O) of this warp; thread layout of the O fragments as in attn_mma.cuh
    if (sub_active) {
#pragma unroll
        for (int half = 0; half < 2; ++half) {
            const int r = gq + half * 8, e = half * 2;
#pragma unroll
            for (int hi = 0; hi < 2; ++hi) {
                *reinterpret_cast<float4*>(mo + r * 64 + 32 * hi + 8 * t) = make_float4(o[4 * hi][e], o[4 * hi + 1][e], o[4 * hi + 2][e], o[4 * hi + 3][e]);
                *reinterpret_cast<float4*>(mo + r * 64 + 32 * hi + 8 * t + 4) = make_float4(o[4 * hi][e + 1], o[4 * hi + 1][e + 1], o[4 * hi + 2][e + 1], o[4 * hi + 3][e + 1]);
            }
        }
        if (t == 0) { mo[1024 + gq] = m0; mo[1024 + gq + 8] = m1; mo[1040 + gq] = l0; mo[1040 + gq + 8] = l1; }
    }
    __syncthreads();
    const long long a4 = probe ? clock64() : 0;
    // merge in warp order, normalise, store as fp16x2: 4 consecutive head dims per thread and round
    for (int idx = tid; idx < nq_sub * 16 * 16; idx += THREADS) {
        const int row = idx >> 4, d4 = idx & 15, sb = row >> 4, lr = row & 15;
        if (row >= nq) continue;
        const float* base = reinterpret_cast<const float*>(smem + OFF_AMG + sb * wps * AMG_WARP);
        float mx = -INFINITY;
        for (int ww = 0; ww < wps; ++ww) mx = fmaxf(mx, base[ww * (AMG_WARP / 4) + 1024 + lr]);
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        float lsum = 0.f;
        for (int ww = 0; ww < wps; ++ww) {
            const float* bw = base + ww * (AMG_WARP / 4);
            const float e = expf(bw[1024 + lr] - mx);             // 0 for a warp that saw no tile (max = -inf)
            const float4 ov = *reinterpret_cast<const float4*>(bw + lr * 64 + d4 * 4);
            acc.x += ov.x * e; acc.y += ov.y * e; acc.z += ov.z * e; acc.w += ov.w * e;
            lsum += bw[1040 + lr] * e;
        }
        const float inv = 1.0f / lsum;
        acc.x *= inv; acc.y *= inv; acc.z *= inv; acc.w *= inv;
        store_split4(Os + (size_t)(o0 + row) * ldo + head * 64 + d4 * 4, os_stride, acc, 2, ovf);
    }
    if (probe) {
        const long long a5 = clock64();
        const int o = Q_SMEM ? 40 : 32;
        prof[o] += a1 - a0; prof[o + 1] += a2 - a1; prof[o + 2] += a3 - a2; prof[o + 3] += a4 - a3; prof[o + 4] += a5 - a4; prof[o + 5] += 1;
    }
}

// One 32 x TNT output tile (TNT = 64 or 32).  FUSE (TNT = 64): phase 5 -- the tile is the cross-attention query of (rows, head n0 / 64);
// it is scaled, placed in the attention routine's Q buffer and consumed in place.
template <bool FUSE, int TNT, int NV>
__device__ __forceinline__ void gemm_item(const Params& p, const GemmDesc& d, uint8_t* smem, int P, int mt, int row0, int nrows, int n0, int wf, int li) {
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5, g = lane >> 2, t = lane & 3;
    const int kg = w >> 2, wm = (w >> 1) & 1, wn = w & 1;  // k group; warp tile: rows 16 wm .. +15, columns (TNT / 2) wn .. + TNT / 2 - 1
    constexpr int NJ = TNT / 16;                           // 8-column MMA tiles per warp
    const uint32_t sbase = s_u32(smem);
    const long long w_split = (long long)d.N * d.K;
    const uint16_t* A = (n0 >= d.n_switch) ? d.A1 : d.A0;

    __syncthreads();                                       // the previous item of this CTA is done with the shared buffers
    const bool probe = p.prof && blockIdx.x == 0 && tid == 0 && !FUSE && d.K == p.E && d.R != nullptr;      // CTA 0, residual projections
    const long long c0 = probe ? clock64() : 0;
    const int KP = d.K / (2 * TK);
    // per-thread source / destination bases: piece c = tid & 7 of row (tid >> 3) & 31; the copies of a pair differ by constants only
    const int lr_ = (tid >> 3) & 31, lc_ = tid & 7;
    const uint16_t* a_src = A + (size_t)(row0 + min(lr_, nrows - 1)) * d.lda + lc_ * 8;
    const uint16_t* w_src = d.W + (size_t)(n0 + lr_) * d.K + lc_ * 8;
    const uint32_t a_dst = lr_ * 128 + ((lc_ ^ (lr_ & 7)) << 4);          // rows r and r + 32 of W share (r & 7)
    const size_t w_row32 = (size_t)32 * d.K;
    auto issue = [&](int pi) {
        const uint32_t sp = sbase + (pi % NPAIR) * PAIR_BYTES;
        const int k0 = pi * 2 * TK;
#pragma unroll
        for (int ch = 0; ch < 2; ++ch) {
#pragma unroll
            for (int sp2 = 0; sp2 < 2; ++sp2) {
                cp_async16(sp + ch * CHUNK_BYTES + sp2 * SPLIT_A + a_dst, a_src + sp2 * d.a_split + k0 + ch * TK, true);
                const uint16_t* ws = w_src + sp2 * w_split + k0 + ch * TK;
                const uint32_t wd = sp + ch * CHUNK_BYTES + CHUNK_A + sp2 * SPLIT_W + a_dst;
                cp_async16(wd, ws, true);
                if (TNT == 64) cp_async16(wd + 32 * 128, ws + w_row32, true);
            }
        }
    };
#pragma unroll
    for (int s = 0; s < NPAIR - 1; ++s) { if (s < KP) issue(s); cp_async_commit(); }
    const long long c1 = probe ? clock64() : 0;
    long long c2 = 0;

    // bias / residual of this thread's outputs: loaded now (they cross the same L2 latency as the operands), consumed in the epilogue.  The
    // residual aliases the output; every element is read and written by the same thread only.
    float2 bs[NJ], rv[NJ][2];
    if (!FUSE && kg == 0) {
#pragma unroll
        for (int j = 0; j < NJ; ++j) {
            const int col = n0 + wn * (TNT / 2) + j * 8 + 2 * t;
            bs[j] = d.bias ? __ldg(reinterpret_cast<const float2*>(d.bias + col)) : make_float2(0.f, 0.f);
#pragma unroll
            for (int half = 0; half < 2; ++half) {
                const int lr = wm * 16 + g + half * 8;
                rv[j][half] = (d.R && lr < nrows) ? __ldcg(reinterpret_cast<const float2*>(d.R + (size_t)(row0 + lr) * d.ldr + col)) : make_float2(0.f, 0.f);
            }
        }
    }

    float tot[NJ][4], mn[NJ][4], cr[NJ][4];
#pragma unroll
    for (int j = 0; j < NJ; ++j)
#pragma unroll
        for (int e = 0; e < 4; ++e) { tot[j][e] = 0.f; mn[j][e] = 0.f; cr[j][e] = 0.f; }

    for (int pi = 0; pi < KP; ++pi) {
        cp_async_wait<NPAIR - 2>();
        __syncthreads();                                   // pair pi has landed for every thread; pair pi - 1 is consumed
        if (probe && pi == 0) c2 = clock64();
        if (pi + NPAIR - 1 < KP) issue(pi + NPAIR - 1);
        cp_async_commit();
        const uint32_t sa = sbase + (pi % NPAIR) * PAIR_BYTES + kg * CHUNK_BYTES, sw = sa + CHUNK_A;
#pragma unroll
        for (int ks = 0; ks < TK / 16; ++ks) {
            const int mi = lane >> 3, rr = lane & 7;
            uint32_t ah[4], al[4];
            {
                const int r = wm * 16 + (mi & 1) * 8 + rr, c = ks * 2 + (mi >> 1);
                const uint32_t addr = sa + r * 128 + ((c ^ (r & 7)) << 4);
                ldsm_x4(addr, ah[0], ah[1], ah[2], ah[3]);
                ldsm_x4(addr + SPLIT_A, al[0], al[1], al[2], al[3]);
            }
#pragma unroll
            for (int jp = 0; jp < NJ / 2; ++jp) {          // pairs of 8-column tiles
                const int n = wn * (TNT / 2) + jp * 16 + (mi >> 1) * 8 + rr, c = ks * 2 + (mi & 1);
                const uint32_t addr = sw + n * 128 + ((c ^ (n & 7)) << 4);
                uint32_t bh[4], bl[4];
                ldsm_x4(addr, bh[0], bh[1], bh[2], bh[3]);
                ldsm_x4(addr + SPLIT_W, bl[0], bl[1], bl[2], bl[3]);
                mma_f16(cr[2 * jp], al, bh[0], bh[1]);     mma_f16(cr[2 * jp + 1], al, bh[2], bh[3]);
                mma_f16(cr[2 * jp], ah, bl[0], bl[1]);     mma_f16(cr[2 * jp + 1], ah, bl[2], bl[3]);
                mma_f16(mn[2 * jp], ah, bh[0], bh[1]);     mma_f16(mn[2 * jp + 1], ah, bh[2], bh[3]);
            }
        }
        if ((pi & 1) || pi == KP - 1) {                    // drain every 128 k of this group: short tensor-core chains, round-to-nearest adds in between
#pragma unroll
            for (int j = 0; j < NJ; ++j)
#pragma unroll
                for (int e = 0; e < 4; ++e) { tot[j][e] += mn[j][e] + cr[j][e]; mn[j][e] = 0.f; cr[j][e] = 0.f; }
        }
    }
    const long long c3 = probe ? clock64() : 0;
    // split-K hand-over: group 1 -> shared memory -> group 0 (fixed order: even chunks + odd chunks)
    float* red = reinterpret_cast<float*>(smem + OFF_RED);
    if (kg == 1) {
#pragma unroll
        for (int j = 0; j < NJ; ++j) *reinterpret_cast<float4*>(red + ((tid - 128) * NJ + j) * 4) = make_float4(tot[j][0], tot[j][1], tot[j][2], tot[j][3]);
    }
    __syncthreads();                                       // (also: every warp is done with the operand ring)
    if (kg == 0) {
#pragma unroll
        for (int j = 0; j < NJ; ++j) {
            const float4 o = *reinterpret_cast<const float4*>(red + (tid * NJ + j) * 4);
            tot[j][0] += o.x; tot[j][1] += o.y; tot[j][2] += o.z; tot[j][3] += o.w;
        }
    }

    if constexpr (FUSE) {
        float* Qs = reinterpret_cast<float*>(smem);
        if (kg == 0) {
#pragma unroll
            for (int j = 0; j < NJ; ++j) {
                const int col = wn * 32 + j * 8 + 2 * t;
                const float b0 = d.bias[n0 + col], b1 = d.bias[n0 + col + 1];
                const int r0 = wm * 16 + g;
                *reinterpret_cast<float2*>(Qs + r0 * AM_SQ + col) = make_float2(fmaf(tot[j][0], d.scale, b0) * 0.125f, fmaf(tot[j][1], d.scale, b1) * 0.125f);
                *reinterpret_cast<float2*>(Qs + (r0 + 8) * AM_SQ + col) = make_float2(fmaf(tot[j][2], d.scale, b0) * 0.125f, fmaf(tot[j][3], d.scale, b1) * 0.125f);
            }
        }
        __syncthreads();
        const size_t ldkv = (size_t)p.Ld * p.E;
        attn_item<true>(smem, Qs, AM_SQ, nrows, 2, p.Kc + (size_t)li * p.E, p.Vc + (size_t)li * p.E, (int)ldkv, reinterpret_cast<const int*>(smem + OFF_WFI)[(MAX_WF + 1) + wf],
                        reinterpret_cast<const int*>(smem + OFF_WFI)[2 * (MAX_WF + 1) + wf], n0 / 64,
                        d.Cs, d.cs_split, d.ldcs, row0, p.state + 4, p.prof);
        return;
    }

    const long long c4 = probe ? clock64() : 0;
    if (kg == 0) {
#pragma unroll
    for (int j = 0; j < NJ; ++j) {
        const int col = n0 + wn * (TNT / 2) + j * 8 + 2 * t;
#pragma unroll
        for (int half = 0; half < 2; ++half) {
            const int lr = wm * 16 + g + half * 8;
            if (lr >= nrows) continue;
            float v0 = fmaf(tot[j][2 * half], d.scale, bs[j].x), v1 = fmaf(tot[j][2 * half + 1], d.scale, bs[j].y);
            if (d.relu) { v0 = fmaxf(v0, 0.f); v1 = fmaxf(v1, 0.f); }
            const size_t row = (size_t)(row0 + lr);
            if (d.C) {
                if (d.R) { v0 = rv[j][half].x + v0; v1 = rv[j][half].y + v1; }
                *reinterpret_cast<float2*>(d.C + row * d.ldc + col) = make_float2(v0, v1);
            }
            if (d.Cs) {
                uint32_t hi, lo;
                split_pair(v0, v1, hi, lo);
                if (!(fmaxf(fabsf(v0), fabsf(v1)) <= 65504.f)) p.state[4] = 1;
                *reinterpret_cast<uint32_t*>(d.Cs + row * d.ldcs + col) = hi;
                *reinterpret_cast<uint32_t*>(d.Cs + d.cs_split + row * d.ldcs + col) = lo;
            }
        }
    }
    }
    const long long c5 = probe ? clock64() : 0;
    if (probe) {
        p.prof[11] += c1 - c0; p.prof[12] += c2 - c1; p.prof[13] += c3 - c2; p.prof[14] += c4 - c3; p.prof[15] += c5 - c4; p.prof[27] += 1;
    }
}

template <bool FUSE, int TNT, int NV>
__device__ __forceinline__ void gemm_phase(const Params& p, const GemmDesc& d, uint8_t* smem, const int* tile_first, int P, int li) {
    const int m_tiles = tile_first[p.N], n_tiles = d.N / TNT;
    for (int item = blockIdx.x; item < m_tiles * n_tiles; item += gridDim.x) {
        const int mt = item / n_tiles, nt = item - mt * n_tiles;
        int wf, row0, nrows;
        tile_rows(p, tile_first, mt, P, wf, row0, nrows);
        gemm_item<FUSE, TNT, NV>(p, d, smem, P, mt, row0, nrows, nt * TNT, wf, li);
    }
}

// self-attention over the whole prefix, no mask: item = (sequence, head, tile of 16 * nq_sub query positions)
__device__ __forceinline__ void self_attn_phase(const Params& p, uint8_t* smem, int P) {
    const int nsub = (P + 15) >> 4;
    const int nq_sub = (P <= 16) ? 1 : (P <= 32) ? 2 : (P <= 64) ? 4 : (P <= 128) ? 2 : 1;    // short prefixes: rows over the warps; long: keys
    const int qt_n = (nsub + nq_sub - 1) / nq_sub;
    const int items = p.B * p.H * qt_n;
    const int E = p.E;
    for (int item = blockIdx.x; item < items; item += gridDim.x) {
        const int qt = item % qt_n, hd = (item / qt_n) % p.H, b = item / (qt_n * p.H);
        const long long k0 = (long long)b * P, q0 = k0 + qt * 16 * nq_sub;
        __syncthreads();                                   // the previous item of this CTA is done with the shared buffers
        attn_item<false>(smem, p.qkv + (size_t)q0 * 3 * E + hd * 64, 3 * E, min(16 * nq_sub, P - qt * 16 * nq_sub), nq_sub, p.qkv + E, p.qkv + 2 * E, 3 * E,
                         k0, P, hd, p.atts, p.ssE, E, q0, p.state + 4, p.prof);
    }
}

// decoder.norm of the last position in float64 + folded pointer head + first-max argmax + append (pointer_kernel<true>'s arithmetic)
template <int NV>
__device__ __forceinline__ void head_phase(const Params& p, uint8_t* smem, int P, int step) {
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5, E = p.E;
    float* hy = reinterpret_cast<float*>(smem);                        // [E]
    float* bestv = reinterpret_cast<float*>(smem + OFF_MISC);
    int* besti = reinterpret_cast<int*>(smem + OFF_MISC + 32);
    const bool probe = p.prof && blockIdx.x == 0 && tid == 0;
    for (int b = blockIdx.x; b < p.B; b += gridDim.x) {
        __syncthreads();
        const long long h0 = probe ? clock64() : 0;
        {   // the row goes to shared memory first: three short float64 passes over it, nothing held in registers
            const float* xr = p.x + ((size_t)b * P + (P - 1)) * E;
            for (int c = tid; c < E; c += THREADS) hy[c] = __ldcg(xr + c);
        }
        __syncthreads();
        const long long h1 = probe ? clock64() : 0;
        {   // all 8 warps: thread t owns elements t, t + 256, ... (E <= 1024: at most 4); two block-wide float64 reductions through shared memory
            double* red64 = reinterpret_cast<double*>(smem + 8192);          // [8] partials (the staged row uses the first E * 4 <= 4096 bytes)
            float xv[4]; float gw[4], gb[4];
            int ne = 0;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const int c = tid + k * THREADS;
                if (c < E) { xv[k] = hy[c]; gw[k] = __ldg(p.dec_nw + c); gb[k] = __ldg(p.dec_nb + c); ne = k + 1; }
            }
            double s = 0.0;
#pragma unroll
            for (int k = 0; k < 4; ++k) if (k < ne) s += (double)xv[k];
            s = e64::warp_sum_d(s);
            if (lane == 0) red64[w] = s;
            __syncthreads();
            double tot = 0.0;
#pragma unroll
            for (int k = 0; k < THREADS / 32; ++k) tot += red64[k];
            const double mean = tot / E;
            double q = 0.0;
#pragma unroll
            for (int k = 0; k < 4; ++k) if (k < ne) { const double dd = (double)xv[k] - mean; q = fma(dd, dd, q); }
            q = e64::warp_sum_d(q);
            __syncthreads();
            if (lane == 0) red64[w] = q;
            __syncthreads();
            double qt = 0.0;
#pragma unroll
            for (int k = 0; k < THREADS / 32; ++k) qt += red64[k];
            const double rstd = 1.0 / sqrt(qt / E + 1e-5);
#pragma unroll
            for (int k = 0; k < 4; ++k) if (k < ne) hy[tid + k * THREADS] = (float)(((double)xv[k] - mean) * rstd * (double)gw[k] + (double)gb[k]);
        }
        __syncthreads();
        const long long h2 = probe ? clock64() : 0;
        const int wf = p.seq_wf[b];
        const int r0 = reinterpret_cast<const int*>(smem + OFF_WFI)[(MAX_WF + 1) + wf], vl = reinterpret_cast<const int*>(smem + OFF_WFI)[2 * (MAX_WF + 1) + wf];
        const int ldm = E + 4;
        float bv = -INFINITY; int bi = 0x7fffffff;
        for (int j0 = w; j0 < vl; j0 += 2 * (THREADS / 32)) {          // two rows in flight per warp (same per-row arithmetic and row order)
            float4 mv[2][NV]; float mb[2];
#pragma unroll
            for (int u = 0; u < 2; ++u) {
                const int j = j0 + u * (THREADS / 32);
                if (j < vl) {
                    const float* mr = p.memW + (size_t)(r0 + j) * ldm;
#pragma unroll
                    for (int i = 0; i < NV; ++i) if (lane * 4 + i * 128 < E) mv[u][i] = *reinterpret_cast<const float4*>(mr + lane * 4 + i * 128);
                    mb[u] = mr[E];
                }
            }
#pragma unroll
            for (int u = 0; u < 2; ++u) {
                const int j = j0 + u * (THREADS / 32);
                if (j >= vl) continue;
                float sd = 0.f;
#pragma unroll
                for (int i = 0; i < NV; ++i) {
                    if (lane * 4 + i * 128 < E) {
                        const float4 pv = *reinterpret_cast<const float4*>(&hy[lane * 4 + i * 128]);
                        sd = fmaf(mv[u][i].x, pv.x, sd); sd = fmaf(mv[u][i].y, pv.y, sd); sd = fmaf(mv[u][i].z, pv.z, sd); sd = fmaf(mv[u][i].w, pv.w, sd);
                    }
                }
                const float s = (float)(warp_sum((double)sd) + (double)mb[u]);
                if (lane == 0) p.logits[(size_t)b * p.Lrows + j] = s;
                if (bi == 0x7fffffff || s > bv || (s != s && bv == bv)) { bv = s; bi = j; }     // first max; like torch.argmax a NaN counts as the maximum
            }
        }
        for (int j = vl + tid; j < p.Lrows; j += THREADS) p.logits[(size_t)b * p.Lrows + j] = -FLT_MAX;   // finfo.min
        const long long h3 = probe ? clock64() : 0;
        if (lane == 0) { bestv[w] = bv; besti[w] = bi; }
        __syncthreads();
        if (tid == 0) {
            float v = bestv[0]; int idx = besti[0];
            for (int k = 1; k < THREADS / 32; ++k) {
                if (besti[k] == 0x7fffffff) continue;
                const bool vn = (v != v), kn = (bestv[k] != bestv[k]);
                if (idx == 0x7fffffff || (kn && (!vn || besti[k] < idx)) || (!vn && !kn && (bestv[k] > v || (bestv[k] == v && besti[k] < idx)))) { v = bestv[k]; idx = besti[k]; }
            }
            p.tok[(size_t)P * p.B + b] = idx;
            // parallel (model_para.py:232): sequences that did NOT emit a special token; seq2seq (model.py:207-210): EOS tokens (config.py:44)
            if (p.mode == 0 ? (idx >= p.num_token) : (idx == 3)) atomicAdd(p.counts + step, 1);
            if (probe) { const long long h4 = clock64(); p.prof[48] += h1 - h0; p.prof[49] += h2 - h1; p.prof[50] += h3 - h2; p.prof[51] += h4 - h3; p.prof[52] += 1; }
        }
    }
}

template <int NV>
__global__ void __launch_bounds__(THREADS, 1) decode_persistent_kernel(const __grid_constant__ Params p) {
    extern __shared__ __align__(128) uint8_t pd_smem[];
    uint8_t* smem = pd_smem;
    int* tile_first = reinterpret_cast<int*>(smem + OFF_TILE);
    {
        int* wfi = reinterpret_cast<int*>(smem + OFF_WFI);
        for (int i = threadIdx.x; i <= p.N; i += THREADS) {
            wfi[i] = p.seq_off[i];
            if (i < p.N) { wfi[(MAX_WF + 1) + i] = p.row_off[i]; wfi[2 * (MAX_WF + 1) + i] = p.vlen[i]; }
        }
    }
    unsigned target = 0;
    const int E = p.E, FF = p.FF;
    int steps = 0, eos_total = 0, stopped = 0;
    for (int step = 0; step < p.max_steps; ++step) {
        const int P = step + 1;
        __syncthreads();
        if (threadIdx.x == 0) {
            int acc = 0;
            const int* so = reinterpret_cast<const int*>(smem + OFF_WFI);
            for (int wf = 0; wf < p.N; ++wf) { tile_first[wf] = acc; acc += ((so[wf + 1] - so[wf]) * P + TM - 1) / TM; }
            tile_first[p.N] = acc;
        }
        __syncthreads();
        // one call site per routine (each is inlined exactly once): the 10 phases of a layer are described by data
        for (int lp = 0; lp < 10 * p.Ld; ++lp) {
            const int li = lp / 10, ph = lp - 10 * li;
            const LayerP& L = p.L[li];
            const long long t0 = clock64();
            if (ph == 0 || ph == 4 || ph == 7) {
                const float* g = (ph == 0) ? L.n1w : (ph == 4) ? L.n2w : L.n3w;
                const float* b = (ph == 0) ? L.n1b : (ph == 4) ? L.n2b : L.n3b;
                ln_phase<NV>(p, P, ph == 0 && li == 0, g, b, (ph == 4) ? nullptr : p.xs, (ph == 7) ? nullptr : p.xps);
            } else if (ph == 2) {
                self_attn_phase(p, smem, P);
            } else {
                GemmDesc d{};
                d.n_switch = 1 << 30; d.lda = E; d.K = E; d.a_split = p.ssE;
                switch (ph) {
                    case 1:     // q, k from LN1(x) + qpos, v from LN1(x)
                        d.A0 = p.xps; d.A1 = p.xs; d.n_switch = 2 * E; d.W = L.w_sa_in; d.N = 3 * E; d.scale = L.s_sa_in; d.bias = L.b_sa_in; d.C = p.qkv; d.ldc = 3 * E;
                        break;
                    case 3:     // self-attention output projection + residual
                        d.A0 = p.atts; d.W = L.w_sa_out; d.N = E; d.scale = L.s_sa_out; d.bias = L.b_sa_out; d.C = p.x; d.ldc = E; d.R = p.x; d.ldr = E;
                        break;
                    case 5:     // cross-attention: query projection fused with the attention core; K / V from the once-per-wireframe cache
                        d.A0 = p.xps; d.W = L.w_ca_q; d.N = E; d.scale = L.s_ca_q; d.bias = L.b_ca_q; d.Cs = p.atts; d.cs_split = p.ssE; d.ldcs = E;
                        break;
                    case 6:     // cross-attention output projection + residual
                        d.A0 = p.atts; d.W = L.w_ca_out; d.N = E; d.scale = L.s_ca_out; d.bias = L.b_ca_out; d.C = p.x; d.ldc = E; d.R = p.x; d.ldr = E;
                        break;
                    case 8:     // feed-forward, first linear + ReLU, straight to the fp16x2 operand of the second linear
                        d.A0 = p.xs; d.W = L.w_l1; d.N = FF; d.scale = L.s_l1; d.bias = L.b_l1; d.relu = 1; d.Cs = p.hs; d.cs_split = p.ssF; d.ldcs = FF;
                        break;
                    default:    // 9: feed-forward, second linear + residual
                        d.A0 = p.hs; d.lda = FF; d.K = FF; d.a_split = p.ssF; d.W = L.w_l2; d.N = E; d.scale = L.s_l2; d.bias = L.b_l2; d.C = p.x; d.ldc = E;
                        d.R = p.x; d.ldr = E;
                        break;
                }
                if (ph == 5) gemm_phase<true, 64, NV>(p, d, smem, tile_first, P, li);
                else gemm_phase<false, 32, NV>(p, d, smem, tile_first, P, li);
            }
            const long long t1 = clock64();
            grid_sync(p.bar, target);
            if (p.prof && blockIdx.x == 0 && threadIdx.x == 0) { p.prof[ph] += t1 - t0; p.prof[16 + ph] += clock64() - t1; }
        }
        const long long th0 = clock64();
        head_phase<NV>(p, smem, P, step);
        const long long th1 = clock64();
        grid_sync(p.bar, target);
        if (p.prof && blockIdx.x == 0 && threadIdx.x == 0) { p.prof[10] += th1 - th0; p.prof[26] += clock64() - th1; }
        ++steps;
        const int c = __ldcg(p.counts + step);
        if (p.mode == 0) { if (c == 0) stopped = 1; }
        else { eos_total += c; if (eos_total == p.B) stopped = 1; }
        if (stopped) break;
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) { p.state[0] = stopped; p.state[1] = steps; p.state[2] = 0; p.state[3] = eos_total; }
}

}  // namespace pd
}  // namespace ffb
