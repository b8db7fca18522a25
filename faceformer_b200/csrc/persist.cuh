// persist.cuh -- the whole greedy decode loop of a SMALL batch as ONE persistent cooperative kernel (sm_100a).
//
// Reference: SurfaceFormer_Parallel.forward_eval (faceformer/models/model_para.py:216-236) and SurfaceFormer.forward_eval
// (faceformer/models/model.py:192-213) over TransformerDecoderLayer.forward_pre (faceformer/transformer.py:235-256), decoder.norm
// (transformer.py:115-116), project + select_next (model_para.py:173-179,225).
//
// Why: with one wireframe per batch (the reference's own test loop, trainer.py:51; BASELINE configs[0]) a decode step is 65 MB of weights
// against a few hundred activation rows.  As ~72 dependent kernel launches it costs 0.7-0.9 ms; the work itself is a few microseconds per
// phase.  This kernel runs ALL steps of the loop in one launch: the grid (2 CTAs per SM, co-resident through a cooperative launch) walks
// the phases of a step and meets at a grid-wide barrier between them; the stop predicate is evaluated on the device after every step.
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
// GEMM items are 32 x 64 output tiles on 256 threads: two groups of four warps split K (even / odd 64-wide chunks) and are summed in a fixed
// order; mma.sync m16n8k16 on fp16x2 split operands (hi*hi + lo*hi + hi*lo), fp32 accumulation, register accumulators drained with
// round-to-nearest adds every 128 k of a group -- the precision class of gemm_tc.cuh.  Every GEMM operand lives in global memory (L2 at
// these sizes) already split, so a chunk goes cp.async -> ldmatrix -> MMA with no conversion; up to 192 KB (a whole K = 512 item) is in
// flight per CTA.  Weights are the pre-scaled fp16x2 arrays of the tensor-core path.  The attention core is attn_mma.cuh's 3xTF32 tile
// routine, two independent work items per CTA (128 threads each, named barriers).
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
constexpr int OFF_MISC = OFF_TILE + (MAX_WF + 1) * 4;      // head phase: best value / index per warp
constexpr int SMEM_BYTES = ((OFF_MISC + 128 + 127) / 128) * 128;
constexpr int ATT_HALF = ((AM_SMEM_BYTES + 1023) / 1024) * 1024;       // attention buffers of one 128-thread half (alias the ring)
static_assert(2 * ATT_HALF <= OFF_RED, "the attention tile buffers alias the operand ring");
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
        do { asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(bar) : "memory"); } while (v < target);
    }
    __syncthreads();
}

// tile t of the step -> (wireframe, first row, rows).  Tiles never straddle wireframes (the fused cross-attention needs one K / V set per tile).
__device__ __forceinline__ void tile_rows(const Params& p, const int* tile_first, int t, int P, int& wf, int& row0, int& nrows) {
    wf = 0;
    while (wf + 1 < p.N && tile_first[wf + 1] <= t) ++wf;
    const int first = p.seq_off[wf] * P, end = p.seq_off[wf + 1] * P;
    row0 = first + (t - tile_first[wf]) * TM;
    nrows = min(TM, end - row0);
}

// LayerNorm of every row (one warp per row; the arithmetic of layernorm_split_kernel) -> fp16x2 operands, optionally + qpos[row % P].
// gather (layer 0): the rows are the target embeddings memory[token] (model_para.py:217-219); they are also written to x.
__device__ __forceinline__ void ln_phase(const Params& p, int P, bool gather, const float* g, const float* b, uint16_t* out_plain, uint16_t* out_pos) {
    const int lane = threadIdx.x & 31, E = p.E, M = p.B * P;
    const int nv = E >> 7;
    for (int r = blockIdx.x * (THREADS / 32) + (threadIdx.x >> 5); r < M; r += gridDim.x * (THREADS / 32)) {
        const int pos = r % P;
        const float* xr;
        if (gather) { const int sq = r / P; xr = p.mem + (size_t)(p.row_off[p.seq_wf[sq]] + __ldcg(p.tok + (size_t)pos * p.B + sq)) * E; }
        else xr = p.x + (size_t)r * E;
        float4 v[8];
        float s = 0.f;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (i < nv) {
                v[i] = __ldcg(reinterpret_cast<const float4*>(xr + i * 128 + lane * 4));
                if (gather) *reinterpret_cast<float4*>(p.x + (size_t)r * E + i * 128 + lane * 4) = v[i];
                s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
            }
        }
        const float mean = warp_sum(s) / (float)E;
        float q = 0.f;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (i < nv) {
                v[i].x -= mean; v[i].y -= mean; v[i].z -= mean; v[i].w -= mean;
                q += (v[i].x * v[i].x + v[i].y * v[i].y) + (v[i].z * v[i].z + v[i].w * v[i].w);
            }
        }
        const float rstd = 1.0f / sqrtf(warp_sum(q) / (float)E + 1e-5f);
        const float* prow = p.qpos + (size_t)pos * E;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (i < nv) {
                const int c = i * 128 + lane * 4;
                const float4 gm = __ldg(reinterpret_cast<const float4*>(g + c)), bt = __ldg(reinterpret_cast<const float4*>(b + c));
                float4 o;
                o.x = v[i].x * rstd * gm.x + bt.x; o.y = v[i].y * rstd * gm.y + bt.y;
                o.z = v[i].z * rstd * gm.z + bt.z; o.w = v[i].w * rstd * gm.w + bt.w;
                if (out_plain) store_split4(out_plain + (size_t)r * E + c, p.ssE, o, 2, p.state + 4);
                if (out_pos) {
                    const float4 pp = __ldg(reinterpret_cast<const float4*>(prow + c));
                    o.x += pp.x; o.y += pp.y; o.z += pp.z; o.w += pp.w;
                    store_split4(out_pos + (size_t)r * E + c, p.ssE, o, 2, p.state + 4);
                }
            }
        }
    }
}

// One 32 x 64 output tile.  FUSE: phase 5 -- the tile is the cross-attention query of (rows, head n0 / 64); it is scaled, placed in the
// attention routine's Q buffer and consumed in place by the first four warps.
template <bool FUSE>
__device__ __forceinline__ void gemm_item(const Params& p, const GemmDesc& d, uint8_t* smem, int row0, int nrows, int n0, int wf, int li) {
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5, g = lane >> 2, t = lane & 3;
    const int kg = w >> 2, wm = (w >> 1) & 1, wn = w & 1;  // k group; warp tile: rows 16 wm .. +15, columns 32 wn .. +31
    const uint32_t sbase = s_u32(smem);
    const long long w_split = (long long)d.N * d.K;
    const uint16_t* A = (n0 >= d.n_switch) ? d.A1 : d.A0;

    __syncthreads();                                       // the previous item of this CTA is done with the shared buffers
    const int KP = d.K / (2 * TK);
    auto issue = [&](int pi) {
        const uint32_t sp = sbase + (pi % NPAIR) * PAIR_BYTES;
#pragma unroll
        for (int i = 0; i < 4; ++i) {                      // A: 2 chunks x 2 splits x 32 rows x 8 pieces of 8 halves
            const int idx = tid + i * THREADS, ch = idx >> 9, sp2 = (idx >> 8) & 1, r = (idx >> 3) & 31, c = idx & 7;
            cp_async16(sp + ch * CHUNK_BYTES + sp2 * SPLIT_A + r * 128 + ((c ^ (r & 7)) << 4),
                       A + sp2 * d.a_split + (size_t)(row0 + min(r, nrows - 1)) * d.lda + (2 * pi + ch) * TK + c * 8, true);
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) {                      // W: 2 chunks x 2 splits x 64 rows x 8 pieces
            const int idx = tid + i * THREADS, ch = idx >> 10, sp2 = (idx >> 9) & 1, n = (idx >> 3) & 63, c = idx & 7;
            cp_async16(sp + ch * CHUNK_BYTES + CHUNK_A + sp2 * SPLIT_W + n * 128 + ((c ^ (n & 7)) << 4),
                       d.W + sp2 * w_split + (size_t)(n0 + n) * d.K + (2 * pi + ch) * TK + c * 8, true);
        }
    };
#pragma unroll
    for (int s = 0; s < NPAIR - 1; ++s) { if (s < KP) issue(s); cp_async_commit(); }

    float tot[4][4], mn[4][4], cr[4][4];
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
        for (int e = 0; e < 4; ++e) { tot[j][e] = 0.f; mn[j][e] = 0.f; cr[j][e] = 0.f; }

    for (int pi = 0; pi < KP; ++pi) {
        cp_async_wait<NPAIR - 2>();
        __syncthreads();                                   // pair pi has landed for every thread; pair pi - 1 is consumed
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
            for (int jp = 0; jp < 2; ++jp) {               // pairs of 8-column tiles
                const int n = wn * 32 + jp * 16 + (mi >> 1) * 8 + rr, c = ks * 2 + (mi & 1);
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
            for (int j = 0; j < 4; ++j)
#pragma unroll
                for (int e = 0; e < 4; ++e) { tot[j][e] += mn[j][e] + cr[j][e]; mn[j][e] = 0.f; cr[j][e] = 0.f; }
        }
    }
    // split-K hand-over: group 1 -> shared memory -> group 0 (fixed order: even chunks + odd chunks)
    float* red = reinterpret_cast<float*>(smem + OFF_RED);
    if (kg == 1) {
#pragma unroll
        for (int j = 0; j < 4; ++j) *reinterpret_cast<float4*>(red + ((tid - 128) * 4 + j) * 4) = make_float4(tot[j][0], tot[j][1], tot[j][2], tot[j][3]);
    }
    __syncthreads();                                       // (also: every warp is done with the operand ring)
    if (kg == 0) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const float4 o = *reinterpret_cast<const float4*>(red + (tid * 4 + j) * 4);
            tot[j][0] += o.x; tot[j][1] += o.y; tot[j][2] += o.z; tot[j][3] += o.w;
        }
    }

    if constexpr (FUSE) {
        if (kg == 0) {
            float* Qs = reinterpret_cast<float*>(smem);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int col = wn * 32 + j * 8 + 2 * t;
                const float b0 = d.bias[n0 + col], b1 = d.bias[n0 + col + 1];
                const int r0 = wm * 16 + g;
                *reinterpret_cast<float2*>(Qs + r0 * AM_SQ + col) = make_float2(fmaf(tot[j][0], d.scale, b0) * 0.125f, fmaf(tot[j][1], d.scale, b1) * 0.125f);
                *reinterpret_cast<float2*>(Qs + (r0 + 8) * AM_SQ + col) = make_float2(fmaf(tot[j][2], d.scale, b0) * 0.125f, fmaf(tot[j][3], d.scale, b1) * 0.125f);
            }
            const size_t ldkv = (size_t)p.Ld * p.E;
            attn_mma_tile(reinterpret_cast<float*>(smem), true, nullptr, 0, p.Kc + (size_t)li * p.E, p.Vc + (size_t)li * p.E, (int)ldkv, nullptr, d.ldcs,
                          d.Cs, d.cs_split, 2, p.state + 4, 0, nrows, p.row_off[wf], p.vlen[wf], row0, n0 / 64, tid, 1);
        }
        return;
    }

    if (kg != 0) return;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const int col = n0 + wn * 32 + j * 8 + 2 * t;
        const float b0 = d.bias ? d.bias[col] : 0.f, b1 = d.bias ? d.bias[col + 1] : 0.f;
#pragma unroll
        for (int half = 0; half < 2; ++half) {
            const int lr = wm * 16 + g + half * 8;
            if (lr >= nrows) continue;
            float v0 = fmaf(tot[j][2 * half], d.scale, b0), v1 = fmaf(tot[j][2 * half + 1], d.scale, b1);
            if (d.relu) { v0 = fmaxf(v0, 0.f); v1 = fmaxf(v1, 0.f); }
            const size_t row = (size_t)(row0 + lr);
            if (d.C) {
                if (d.R) {
                    const float2 rv = __ldcg(reinterpret_cast<const float2*>(d.R + row * d.ldr + col));
                    v0 = rv.x + v0; v1 = rv.y + v1;
                }
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

template <bool FUSE>
__device__ __forceinline__ void gemm_phase(const Params& p, const GemmDesc& d, uint8_t* smem, const int* tile_first, int P, int li) {
    const int m_tiles = tile_first[p.N], n_tiles = d.N / TN;
    for (int item = blockIdx.x; item < m_tiles * n_tiles; item += gridDim.x) {
        const int mt = item / n_tiles, nt = item - mt * n_tiles;
        int wf, row0, nrows;
        tile_rows(p, tile_first, mt, P, wf, row0, nrows);
        gemm_item<FUSE>(p, d, smem, row0, nrows, nt * TN, wf, li);
    }
}

// self-attention over the whole prefix, no mask: item = (sequence, head, tile of 64 query positions); two items per CTA at a time
__device__ __forceinline__ void self_attn_phase(const Params& p, uint8_t* smem, int P) {
    const int qt_n = (P + AM_BQ - 1) / AM_BQ;
    const int items = p.B * p.H * qt_n;
    const int E = p.E, half = threadIdx.x >> 7;
    for (int item = blockIdx.x * 2 + half; item < items; item += gridDim.x * 2) {
        const int qt = item % qt_n, hd = (item / qt_n) % p.H, b = item / (qt_n * p.H);
        const long long k0 = (long long)b * P, q0 = k0 + qt * AM_BQ;
        attn_mma_tile(reinterpret_cast<float*>(smem + half * ATT_HALF), false, p.qkv, 3 * E, p.qkv + E, p.qkv + 2 * E, 3 * E, nullptr, E, p.atts, p.ssE, 2,
                      p.state + 4, q0, min(AM_BQ, P - qt * AM_BQ), k0, P, q0, hd, threadIdx.x & 127, 1 + half);
    }
}

// decoder.norm of the last position in float64 + folded pointer head + first-max argmax + append (pointer_kernel<true>'s arithmetic)
__device__ __forceinline__ void head_phase(const Params& p, uint8_t* smem, int P, int step) {
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5, E = p.E;
    float* hy = reinterpret_cast<float*>(smem);                        // [E]
    float* bestv = reinterpret_cast<float*>(smem + OFF_MISC);
    int* besti = reinterpret_cast<int*>(smem + OFF_MISC + 32);
    for (int b = blockIdx.x; b < p.B; b += gridDim.x) {
        __syncthreads();
        if (w == 0) {
            const float* xr = p.x + ((size_t)b * P + (P - 1)) * E;
            double v[32];
            const int n = E >> 5;
            double s = 0.0;
#pragma unroll
            for (int i = 0; i < 32; ++i) if (i < n) { v[i] = (double)__ldcg(xr + lane + 32 * i); s += v[i]; }
            const double mean = e64::warp_sum_d(s) / E;
            double q = 0.0;
#pragma unroll
            for (int i = 0; i < 32; ++i) if (i < n) { const double dd = v[i] - mean; q = fma(dd, dd, q); }
            const double rstd = 1.0 / sqrt(e64::warp_sum_d(q) / E + 1e-5);
#pragma unroll
            for (int i = 0; i < 32; ++i)
                if (i < n) { const int c = lane + 32 * i; hy[c] = (float)((v[i] - mean) * rstd * (double)p.dec_nw[c] + (double)p.dec_nb[c]); }
        }
        __syncthreads();
        const int wf = p.seq_wf[b];
        const int r0 = p.row_off[wf], vl = p.vlen[wf];
        const int ldm = E + 4;
        float bv = -INFINITY; int bi = 0x7fffffff;
        for (int j = w; j < vl; j += THREADS / 32) {
            const float* mr = p.memW + (size_t)(r0 + j) * ldm;
            float sd = 0.f;
            for (int c = lane * 4; c < E; c += 128) {
                const float4 mv = *reinterpret_cast<const float4*>(mr + c);
                const float4 pv = *reinterpret_cast<const float4*>(&hy[c]);
                sd = fmaf(mv.x, pv.x, sd); sd = fmaf(mv.y, pv.y, sd); sd = fmaf(mv.z, pv.z, sd); sd = fmaf(mv.w, pv.w, sd);
            }
            const float s = (float)(warp_sum((double)sd) + (double)mr[E]);
            if (lane == 0) p.logits[(size_t)b * p.Lrows + j] = s;
            if (bi == 0x7fffffff || s > bv || (s != s && bv == bv)) { bv = s; bi = j; }     // first max; like torch.argmax a NaN counts as the maximum
        }
        for (int j = vl + tid; j < p.Lrows; j += THREADS) p.logits[(size_t)b * p.Lrows + j] = -FLT_MAX;   // finfo.min
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
        }
    }
}

__global__ void __launch_bounds__(THREADS, 1) decode_persistent_kernel(const __grid_constant__ Params p) {
    extern __shared__ __align__(128) uint8_t pd_smem[];
    uint8_t* smem = pd_smem;
    int* tile_first = reinterpret_cast<int*>(smem + OFF_TILE);
    unsigned target = 0;
    const int E = p.E, FF = p.FF;
    int steps = 0, eos_total = 0, stopped = 0;
    for (int step = 0; step < p.T - 1; ++step) {
        const int P = step + 1;
        __syncthreads();
        if (threadIdx.x == 0) {
            int acc = 0;
            for (int wf = 0; wf < p.N; ++wf) { tile_first[wf] = acc; acc += ((p.seq_off[wf + 1] - p.seq_off[wf]) * P + TM - 1) / TM; }
            tile_first[p.N] = acc;
        }
        __syncthreads();
        // one call site per routine (each is inlined exactly once): the 10 phases of a layer are described by data
        for (int lp = 0; lp < 10 * p.Ld; ++lp) {
            const int li = lp / 10, ph = lp - 10 * li;
            const LayerP& L = p.L[li];
            GemmDesc d{};
            d.n_switch = 1 << 30; d.lda = E; d.K = E; d.a_split = p.ssE;
            const long long t0 = clock64();
            switch (ph) {
                case 0: ln_phase(p, P, li == 0, L.n1w, L.n1b, p.xs, p.xps); break;
                case 4: ln_phase(p, P, false, L.n2w, L.n2b, nullptr, p.xps); break;
                case 7: ln_phase(p, P, false, L.n3w, L.n3b, p.xs, nullptr); break;
                case 2: self_attn_phase(p, smem, P); break;
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
            if (ph == 5) gemm_phase<true>(p, d, smem, tile_first, P, li);
            else if (ph == 1 || ph == 3 || ph == 6 || ph >= 8) gemm_phase<false>(p, d, smem, tile_first, P, li);
            const long long t1 = clock64();
            grid_sync(p.bar, target);
            if (p.prof && blockIdx.x == 0 && threadIdx.x == 0) { p.prof[ph] += t1 - t0; p.prof[16 + ph] += clock64() - t1; }
        }
        const long long th0 = clock64();
        head_phase(p, smem, P, step);
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
