// enc64.cuh -- float64 kernels for the two stages whose fp32 rounding dominates the pointer-logit error budget
// (tests/golden/noise_budget.npz, profiles/logit_noise.py: in the reference's own fp32 run the encoder memory contributes
// 4.6-4.9e-5 of its 4.2-7.2e-5 distance from the float64 result, project + pointer dot 1.2-3.8e-5, the whole decoder stack
// 1.5-2.8e-5): the ENCODER (embedding.py:23-38, transformer.py:70-83,164-176) with the once-per-wireframe cross-attention
// K / V projections (transformer.py:248-251), and the pointer HEAD (decoder.norm, transformer.py:115-116; project,
// model_para.py:225; select_next's dot product, model_para.py:173-177).  Both are ~1 % of the algorithmic FLOPs of a decode
// (SURVEY.md 8d), so running them on the FP64 pipe costs a few per cent of a step and removes their share of the error.
// Weights stay fp32 in memory (exact in float64); results are rounded to fp32 once, where the fp32 decode consumes them.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "kernels.cuh"

namespace ffb {
namespace e64 {

__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ double warp_max_d(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// ---- C[M,N] = act(A[M,K] . W[N,K]^T + bias) (+ R): fp64 accumulate, 128 x 64 tile, 8 x 4 outputs per thread -----------------------
constexpr int GBM = 128, GBN = 64, GBK = 16;
struct GemmArgs {
    const void* A; int lda; int a_f32;            // A is double (a_f32 = 0) or float (1)
    const int* a_rows;                            // optional row gather for A
    const float* W; int ldw; const float* bias;   // fp32 parameters
    double* C; int ldc; const int* c_rows;        // fp64 output (may be null) with optional row scatter
    float* C32; int ldc32;                        // fp32 copy of the output (may be null)
    const double* R; int ldr;                     // residual (may alias C)
    int M, N, K, relu;
    const int* stop;
};

// BM x BN output tile, TMR x TNR outputs per thread (256 threads = (BM / TMR) x (BN / TNR)).  <128, 64, 8, 4> is the throughput shape;
// <32, 32, 2, 2> is for the few hundred memory rows of a single wireframe (one 128-row tile would leave all but N / 64 SMs idle and spend
// ~100 us per launch at the FP64 rate of one SM).  Every output is one fma chain over k in ascending order in either shape: bit-identical.
template <int BM, int BN, int TMR, int TNR>
__global__ void __launch_bounds__(256, 2) dgemm_kernel(const GemmArgs a) {
    static_assert((BM / TMR) * (BN / TNR) == 256 && BN / TNR == 16 && TMR % 2 == 0 && TNR % 2 == 0, "thread layout");
    constexpr int EA = BM * GBK / 256, EW = BN * GBK / 256;       // consecutive k per thread of the A / W loaders
    if (a.stop != nullptr && *a.stop != 0) return;
    __shared__ __align__(16) double As[GBK][BM + 2];
    __shared__ __align__(16) double Ws[GBK][BN + 2];
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;
    double acc[TMR][TNR];
#pragma unroll
    for (int i = 0; i < TMR; ++i)
#pragma unroll
        for (int j = 0; j < TNR; ++j) acc[i][j] = 0.0;
    // loaders: A tile BM x 16 -> thread (row = tid / (16 / EA), EA consecutive k); W tile BN x 16 likewise
    const int ar = tid / (GBK / EA), ak = (tid % (GBK / EA)) * EA;
    const int wr = tid / (GBK / EW), wk = (tid % (GBK / EW)) * EW;
    const int arow = m0 + ar;
    long long asrc = -1;
    if (arow < a.M) asrc = a.a_rows ? a.a_rows[arow] : arow;
    const int wrow = n0 + wr;
    // software pipeline: the global loads of k-tile t+1 are in flight while tile t is multiplied out of shared memory
    double av[EA], wv[EW];
    auto load_tile = [&](int k0) {
#pragma unroll
        for (int i = 0; i < EA; ++i) av[i] = 0.0;
#pragma unroll
        for (int i = 0; i < EW; ++i) wv[i] = 0.0;
        if (asrc >= 0) {
            if (a.a_f32) {
                const float* p = reinterpret_cast<const float*>(a.A) + asrc * a.lda + k0 + ak;
                if (EA == 8 && k0 + ak + 8 <= a.K && (a.lda & 3) == 0 && ((uintptr_t)p & 15) == 0) {
                    const float4 u = *reinterpret_cast<const float4*>(p), v = *reinterpret_cast<const float4*>(p + 4);
                    const float t8[8] = {u.x, u.y, u.z, u.w, v.x, v.y, v.z, v.w};
#pragma unroll
                    for (int i = 0; i < EA; ++i) av[i] = t8[i < 8 ? i : 0];
                } else if (EA == 2 && k0 + ak + 2 <= a.K && (a.lda & 1) == 0 && ((uintptr_t)p & 7) == 0) {
                    const float2 u = *reinterpret_cast<const float2*>(p);
                    av[0] = u.x; av[EA - 1] = u.y;
                } else {
#pragma unroll
                    for (int i = 0; i < EA; ++i) if (k0 + ak + i < a.K) av[i] = (double)p[i];
                }
            } else {
                const double* p = reinterpret_cast<const double*>(a.A) + asrc * a.lda + k0 + ak;
                if (k0 + ak + EA <= a.K && (a.lda & 1) == 0 && ((uintptr_t)p & 15) == 0) {
#pragma unroll
                    for (int i = 0; i < EA / 2; ++i) { const double2 u = *reinterpret_cast<const double2*>(p + 2 * i); av[2 * i] = u.x; av[2 * i + 1] = u.y; }
                } else {
#pragma unroll
                    for (int i = 0; i < EA; ++i) if (k0 + ak + i < a.K) av[i] = p[i];
                }
            }
        }
        if (wrow < a.N) {
            const float* p = a.W + (size_t)wrow * a.ldw + k0 + wk;
            if (EW == 4 && k0 + wk + 4 <= a.K && (a.ldw & 3) == 0 && ((uintptr_t)p & 15) == 0) {
                const float4 u = *reinterpret_cast<const float4*>(p);
                const float t4[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
                for (int i = 0; i < EW; ++i) wv[i] = t4[i < 4 ? i : 0];
            } else if (EW == 2 && k0 + wk + 2 <= a.K && (a.ldw & 1) == 0 && ((uintptr_t)p & 7) == 0) {
                const float2 u = *reinterpret_cast<const float2*>(p);
                wv[0] = u.x; wv[EW - 1] = u.y;
            } else {
#pragma unroll
                for (int i = 0; i < EW; ++i) if (k0 + wk + i < a.K) wv[i] = (double)p[i];
            }
        }
    };
    load_tile(0);
    for (int k0 = 0; k0 < a.K; k0 += GBK) {
        __syncthreads();
#pragma unroll
        for (int i = 0; i < EA; ++i) As[ak + i][ar] = av[i];
#pragma unroll
        for (int i = 0; i < EW; ++i) Ws[wk + i][wr] = wv[i];
        __syncthreads();
        if (k0 + GBK < a.K) load_tile(k0 + GBK);
#pragma unroll
        for (int kk = 0; kk < GBK; ++kk) {
            double am[TMR], wn[TNR];
#pragma unroll
            for (int i = 0; i < TMR / 2; ++i) { const double2 u = *reinterpret_cast<const double2*>(&As[kk][ty * TMR + 2 * i]); am[2 * i] = u.x; am[2 * i + 1] = u.y; }
#pragma unroll
            for (int j = 0; j < TNR / 2; ++j) { const double2 u = *reinterpret_cast<const double2*>(&Ws[kk][tx * TNR + 2 * j]); wn[2 * j] = u.x; wn[2 * j + 1] = u.y; }
#pragma unroll
            for (int i = 0; i < TMR; ++i)
#pragma unroll
                for (int j = 0; j < TNR; ++j) acc[i][j] = fma(am[i], wn[j], acc[i][j]);
        }
    }
#pragma unroll
    for (int i = 0; i < TMR; ++i) {
        const int r = m0 + ty * TMR + i;
        if (r >= a.M) continue;
        const long long cr = a.c_rows ? a.c_rows[r] : r;
#pragma unroll
        for (int j = 0; j < TNR; ++j) {
            const int c = n0 + tx * TNR + j;
            if (c >= a.N) continue;
            double v = acc[i][j];
            if (a.bias) v += (double)a.bias[c];
            if (a.relu) v = fmax(v, 0.0);
            if (a.R) v = a.R[cr * a.ldr + c] + v;
            if (a.C) a.C[cr * a.ldc + c] = v;
            if (a.C32) a.C32[cr * a.ldc32 + c] = (float)v;
        }
    }
}

// ---- LayerNorm (eps 1e-5), one warp per row.  x: double, or float with an optional row gather (row = r * in_mul + in_off).
//      y = LN(x); yp = LN(x) + pos[pos_idx ? pos_idx[r] : r % pos_mod] (optional); y32 = fp32 copy (optional) ------------------------
__global__ void __launch_bounds__(256) layernorm64_kernel(const double* __restrict__ x64, const float* __restrict__ x32, int in_mul, int in_off,
                                                          const float* __restrict__ gamma, const float* __restrict__ beta,
                                                          double* __restrict__ y, double* __restrict__ yp, float* __restrict__ y32,
                                                          const float* __restrict__ pos, const int* __restrict__ pos_idx, int pos_mod,
                                                          int M, int E, const int* stop) {
    if (stop != nullptr && *stop != 0) return;
    const int r = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (r >= M) return;
    const size_t src = (size_t)r * in_mul + in_off;
    double v[32];                                   // E <= 1024
    const int n = E >> 5;
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < 32; ++i) {
        if (i < n) {
            const int c = lane + 32 * i;
            v[i] = x64 ? x64[src * E + c] : (double)x32[src * E + c];
            s += v[i];
        }
    }
    const double mean = warp_sum_d(s) / E;
    double q = 0.0;
#pragma unroll
    for (int i = 0; i < 32; ++i) if (i < n) { const double d = v[i] - mean; q = fma(d, d, q); }
    const double rstd = 1.0 / sqrt(warp_sum_d(q) / E + 1e-5);
    const float* prow = pos ? pos + (size_t)(pos_idx ? pos_idx[r] : (r % pos_mod)) * E : nullptr;
#pragma unroll
    for (int i = 0; i < 32; ++i) {
        if (i < n) {
            const int c = lane + 32 * i;
            const double o = (v[i] - mean) * rstd * (double)gamma[c] + (double)beta[c];
            if (y) y[(size_t)r * E + c] = o;
            if (yp) yp[(size_t)r * E + c] = o + (double)prow[c];
            if (y32) y32[(size_t)r * E + c] = (float)o;
        }
    }
}

// out[r] = in[r] + pos[pos_idx[r]]  (memory + pos for the cross-attention keys)
__global__ void add_pos64_kernel(const double* __restrict__ in, const float* __restrict__ pos, const int* __restrict__ pos_idx,
                                 double* __restrict__ out, int M, int E) {
    const long long total = (long long)M * E;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int c = (int)(i % E); const long long r = i / E;
        out[i] = in[i] + (double)pos[(size_t)pos_idx[r] * E + c];
    }
}

// rows 0..3 of every wireframe's packed block <- special-token table (embedding.py:30-32,36)
__global__ void token_rows64_kernel(const float* __restrict__ table, const int* __restrict__ row_off, double* __restrict__ x,
                                    int n_wf, int num_token, int E) {
    const int total = n_wf * num_token * E;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        const int c = i % E, t = (i / E) % num_token, w = i / (E * num_token);
        x[(size_t)(row_off[w] + t) * E + c] = (double)table[t * E + c];
    }
}

// ---- self-attention of the encoder: softmax(q k^T / 8) v per (wireframe, head) ------------------------------------------------------
// qkv: [R, 3E] doubles (q at column 0, k at E, v at 2E); out [R, E].  grid (ceil(max_vlen / A64_ROWS), H, N), 32 * A64_ROWS threads: one
// warp per query row; K / V are staged 32 keys at a time in shared memory for all the CTA's rows (coalesced loads; the per-lane key
// rows are padded to 65 doubles); the row's scores live in shared memory.  dynamic smem: A64_ROWS * (max_vlen + 64) doubles.
constexpr int A64_ROWS = 8;
__global__ void __launch_bounds__(32 * A64_ROWS) attn64_kernel(const double* __restrict__ qkv, double* __restrict__ out,
                                                               const int* __restrict__ row_off, const int* __restrict__ v_len, int E, int max_vlen) {
    extern __shared__ __align__(16) double a64_smem[];
    __shared__ __align__(16) double kv_tile[32][65];
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int wf = blockIdx.z, hd = blockIdx.y;
    const int vl = v_len[wf], r0 = row_off[wf];
    if ((int)blockIdx.x * A64_ROWS >= vl) return;
    const int qi = blockIdx.x * A64_ROWS + w;
    const bool live = qi < vl;
    double* sc = a64_smem + (size_t)w * (max_vlen + 64);
    double* qs = sc + max_vlen;
    const int ld = 3 * E;
    if (live) {
        const double* qrow = qkv + (size_t)(r0 + qi) * ld + hd * 64;
        qs[lane] = qrow[lane] * 0.125; qs[lane + 32] = qrow[lane + 32] * 0.125;      // q * sqrt(1/64) (functional.py:6632), exact
    }
    double m = -INFINITY;
    for (int j0 = 0; j0 < vl; j0 += 32) {
        __syncthreads();
        for (int i = threadIdx.x; i < 32 * 64; i += 32 * A64_ROWS) {
            const int jj = i >> 6, d = i & 63;
            kv_tile[jj][d] = (j0 + jj < vl) ? qkv[(size_t)(r0 + j0 + jj) * ld + E + hd * 64 + d] : 0.0;
        }
        __syncthreads();
        if (live && j0 + lane < vl) {
            double s = 0.0;
#pragma unroll 16
            for (int d = 0; d < 64; ++d) s = fma(qs[d], kv_tile[lane][d], s);
            sc[j0 + lane] = s;
            m = fmax(m, s);
        }
    }
    m = warp_max_d(m);
    double l = 0.0;
    if (live) for (int j = lane; j < vl; j += 32) { const double p = exp(sc[j] - m); sc[j] = p; l += p; }
    l = warp_sum_d(l);
    double a0 = 0.0, a1 = 0.0;
    for (int j0 = 0; j0 < vl; j0 += 32) {
        __syncthreads();
        for (int i = threadIdx.x; i < 32 * 64; i += 32 * A64_ROWS) {
            const int jj = i >> 6, d = i & 63;
            kv_tile[jj][d] = (j0 + jj < vl) ? qkv[(size_t)(r0 + j0 + jj) * ld + 2 * E + hd * 64 + d] : 0.0;
        }
        __syncthreads();
        if (live) {
            const int n = min(32, vl - j0);
            for (int jj = 0; jj < n; ++jj) {
                const double p = sc[j0 + jj];
                a0 = fma(p, kv_tile[jj][lane], a0); a1 = fma(p, kv_tile[jj][lane + 32], a1);
            }
        }
    }
    if (live) {
        double* orow = out + (size_t)(r0 + qi) * E + hd * 64;
        orow[lane] = a0 / l; orow[lane + 32] = a1 / l;
    }
}

}  // namespace e64
}  // namespace ffb
