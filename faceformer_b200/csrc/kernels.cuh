// kernels.cuh -- hand-written sm_100a kernels of the FaceFormer greedy pointer-decode path.
//
// The fp32 SIMT kernels of this file accumulate in fp32 with fp32 operands (the parity contract - token-exact vs the
// reference's CPU fp32 path, pointer logits at fp32-noise level - leaves no room for a plain 16-bit pass, SURVEY.md
// section 7); they serve small steps / batches and odd geometries.  Large steps run on the split-precision tensor-core
// kernels (gemm_tc.cuh, attn_x.cuh, attn_h.cuh), for which this file provides the operand formatting (LayerNorm -> fp16x2 /
// bf16x3 splits), the pointer head, the loop control and the pre / post steps (featurize, parse_faces).
//
// Reference operations each kernel replaces (paths relative to /root/reference):
//   linear_kernel      nn.Linear call sites: transformer.py:134,136,195,197, model_para.py:46,
//                      embedding.py:15,17 and the in/out projections of nn.MultiheadAttention
//                      (torch functional.py:5866-5873,6653); fuses with_pos_embed (transformer.py:
//                      144-145,205-206), bias, ReLU and the residual add (transformer.py:172,176,246,..).
//   layernorm_kernel   nn.LayerNorm (transformer.py:138-139,199-201, model_para.py:37,43)
//   attn_rows_kernel   decoder self-attention core (transformer.py:244-245; NO tgt_mask in eval)
//   attn_tiled_kernel  encoder self-attention / decoder cross-attention core with key-padding mask
//                      realised as "only valid rows exist" (transformer.py:169-171,247-251)
//   gather_tgt_kernel  torch.gather(memory, 0, predicts...) (model_para.py:217-219)
//   pointer_kernel     select_next: bmm + masked_fill + argmax (model_para.py:173-179) + append +
//                      stop predicate (model_para.py:229-233 / model.py:205-210)
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <float.h>
#include <math.h>

namespace ffb {

#define FFB_STOP_CHECK(stop) do { if ((stop) != nullptr && *(stop) != 0) return; } while (0)

// Programmatic dependent launch (decode-step kernels, FFB_OPT_PDL): let the next kernel of the stream be scheduled onto SMs as this
// grid's CTAs retire (its prologue - barrier init, TMEM allocation, descriptor fetch - then overlaps this grid's tail), and wait
// until the previous grid has completed and flushed before touching any memory.  Both are no-ops in a normal launch.
#define FFB_PDL_SYNC() do { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); asm volatile("griddepcontrol.wait;" ::: "memory"); } while (0)

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// bf16x3 split of an fp32 value (operand format of the tensor-core GEMM, gemm_tc.cuh): x = b0 + b1 + b2
__device__ __forceinline__ void split3_bf16(float x, __nv_bfloat16& b0, __nv_bfloat16& b1, __nv_bfloat16& b2) {
    b0 = __float2bfloat16_rn(x);
    const float r1 = x - __bfloat162float(b0);
    b1 = __float2bfloat16_rn(r1);
    const float r2 = r1 - __bfloat162float(b1);
    b2 = __float2bfloat16_rn(r2);
}
// Operand formats of the tensor-core GEMM (gemm_tc.cuh).  fmt 3: bf16x3 (x = b0+b1+b2); fmt 2: fp16x2 (x = h+l).
// Arrays are opaque 16-bit words; split s lives `stride` elements after split s-1.
// fp16 has a narrow range: |x| > 65504 raises *ovf (the host then re-runs the decode in the bf16x3 format).
__device__ __forceinline__ void store_split4(uint16_t* dst, long long stride, float4 v, int fmt, int* ovf) {
    if (fmt == 3) {
        __align__(8) __nv_bfloat16 o0[4], o1[4], o2[4];
        split3_bf16(v.x, o0[0], o1[0], o2[0]); split3_bf16(v.y, o0[1], o1[1], o2[1]);
        split3_bf16(v.z, o0[2], o1[2], o2[2]); split3_bf16(v.w, o0[3], o1[3], o2[3]);
        *reinterpret_cast<uint2*>(dst) = *reinterpret_cast<const uint2*>(o0);
        *reinterpret_cast<uint2*>(dst + stride) = *reinterpret_cast<const uint2*>(o1);
        *reinterpret_cast<uint2*>(dst + 2 * stride) = *reinterpret_cast<const uint2*>(o2);
    } else {
        // packed conversions (one F2FP per pair) and one max-|x| test instead of four compares
        const __half2 h01 = __floats2half2_rn(v.x, v.y), h23 = __floats2half2_rn(v.z, v.w);
        const float2 f01 = __half22float2(h01), f23 = __half22float2(h23);
        const __half2 l01 = __floats2half2_rn(v.x - f01.x, v.y - f01.y), l23 = __floats2half2_rn(v.z - f23.x, v.w - f23.y);
        const bool bad = !(fmaxf(fmaxf(fabsf(v.x), fabsf(v.y)), fmaxf(fabsf(v.z), fabsf(v.w))) <= 65504.f) ||
                         (v.x != v.x) || (v.y != v.y) || (v.z != v.z) || (v.w != v.w);
        *reinterpret_cast<uint2*>(dst) = make_uint2(*reinterpret_cast<const uint32_t*>(&h01), *reinterpret_cast<const uint32_t*>(&h23));
        *reinterpret_cast<uint2*>(dst + stride) = make_uint2(*reinterpret_cast<const uint32_t*>(&l01), *reinterpret_cast<const uint32_t*>(&l23));
        if (bad && ovf) *ovf = 1;
    }
}
__device__ __forceinline__ void store_split1(uint16_t* dst, long long stride, float x, int fmt, int* ovf) {
    if (fmt == 3) {
        __nv_bfloat16 b0, b1, b2;
        split3_bf16(x, b0, b1, b2);
        dst[0] = __bfloat16_as_ushort(b0); dst[stride] = __bfloat16_as_ushort(b1); dst[2 * stride] = __bfloat16_as_ushort(b2);
    } else {
        const __half h = __float2half_rn(x);
        const __half l = __float2half_rn(x - __half2float(h));
        dst[0] = __half_as_ushort(h); dst[stride] = __half_as_ushort(l);
        if (!(fabsf(x) <= 65504.f) && ovf) *ovf = 1;
    }
}

// dst[fmt][n] 16-bit <- split(src[n] * scale)  (weights, once at load time; scale is a power of two)
__global__ void split_array_kernel(const float* __restrict__ src, uint16_t* __restrict__ dst, long long n4, float scale, int fmt,
                                   int* ovf = nullptr) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
        float4 v = reinterpret_cast<const float4*>(src)[i];
        v.x *= scale; v.y *= scale; v.z *= scale; v.w *= scale;
        store_split4(dst + 4 * i, 4 * n4, v, fmt, ovf);
    }
}

// out[0] = max |src[i]| (as int bits of a non-negative float; out must be zeroed first)
__global__ void absmax_kernel(const float* __restrict__ src, long long n, int* __restrict__ out) {
    float m = 0.f;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        m = fmaxf(m, fabsf(src[i]));
    m = warp_max(m);
    if ((threadIdx.x & 31) == 0) atomicMax(out, __float_as_int(m));
}

// benchmark helper: pseudo-random halves; elements of the first split ~ U(-2, 2), later splits ~ 2^-11 of that (like real lo parts)
__global__ void fill_random_half_kernel(uint16_t* __restrict__ dst, long long n, long long split_stride, unsigned seed) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        unsigned x = (unsigned)i * 2654435761u + seed * 40503u; x ^= x >> 15; x *= 2246822519u; x ^= x >> 13;
        const float u = ((x & 0xffffu) / 32768.0f - 1.0f) * 2.0f;
        dst[i] = __half_as_ushort(__float2half_rn(i < split_stride ? u : u * 4.8828125e-4f));
    }
}

// dst[i] = sum of the splits of element i  (test hook: re-sum a split array)
__global__ void sum_split_kernel(const uint16_t* __restrict__ src, long long stride, float* __restrict__ dst, long long n, int fmt) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        if (fmt == 3)
            dst[i] = (__bfloat162float(__ushort_as_bfloat16(src[i])) + __bfloat162float(__ushort_as_bfloat16(src[i + stride]))) +
                     __bfloat162float(__ushort_as_bfloat16(src[i + 2 * stride]));
        else
            dst[i] = __half2float(__ushort_as_half(src[i])) + __half2float(__ushort_as_half(src[i + stride]));
    }
}

// ------------------------------------------------------------------------------------------------
// linear: C[M,N] = act( (A[+pos]) W^T + bias ) (+ R)
// ------------------------------------------------------------------------------------------------
struct LinearArgs {
    const float* A; int lda;          // [M,K] row-major (row stride lda)
    const int* a_rows;                // optional gather: row r of the GEMM reads A[a_rows[r]]
    const float* W; int ldw;          // [N,K] row-major == torch Linear.weight
    const float* bias;                // [N] or null
    float* C; int ldc;                // [M,N]
    const int* c_rows;                // optional scatter: row r is written to C[c_rows[r]]
    const float* R; int ldr;          // residual rows (same row mapping as C) or null; may alias C
    const float* pos; int ldpos;      // positional table; added to A for column tiles n0 < pos_cols
    const int* pos_idx; int pos_mod;  // table row of GEMM row r: pos_mod > 0 ? r % pos_mod : pos_idx[r]
    int pos_cols;                     // multiple of 128 (or >= N)
    int M, N, K;                      // K multiple of 4
    int relu;
    const int* stop;
};

constexpr int LBM = 128, LBN = 128, LBK = 16, LPAD = 4;

__global__ void __launch_bounds__(256, 2) linear_kernel(const LinearArgs a) {
    FFB_STOP_CHECK(a.stop);
    __shared__ __align__(16) float As[2][LBK][LBM + LPAD];
    __shared__ __align__(16) float Bs[2][LBK][LBN + LPAD];

    const int tid = threadIdx.x;
    const int tx = tid & 15, ty = tid >> 4;
    const int m0 = blockIdx.x * LBM, n0 = blockIdx.y * LBN;
    const bool add_pos = (a.pos != nullptr) && (n0 < a.pos_cols);

    // global->smem staging: float4 index f = tid + i*256 -> tile row f>>2, k-quad f&3
    const float* ap[2]; const float* pp[2]; const float* wp[2];
    int srow[2];
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        const int f = tid + i * 256;
        const int row = f >> 2, kq = f & 3;
        srow[i] = row;
        const int gr = m0 + row;
        ap[i] = nullptr; pp[i] = nullptr; wp[i] = nullptr;
        if (gr < a.M) {
            const int src = a.a_rows ? a.a_rows[gr] : gr;
            ap[i] = a.A + (size_t)src * a.lda + kq * 4;
            if (add_pos) {
                const int pi = a.pos_mod > 0 ? (gr % a.pos_mod) : a.pos_idx[gr];
                pp[i] = a.pos + (size_t)pi * a.ldpos + kq * 4;
            }
        }
        const int gn = n0 + row;
        if (gn < a.N) wp[i] = a.W + (size_t)gn * a.ldw + kq * 4;
    }
    const int kq4 = (tid & 3) * 4;

    float acc[8][8];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

    float4 ra[2], rb[2];
    auto gload = [&](int k0) {
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            const bool kv = (k0 + kq4 + 4 <= a.K);
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (ap[i] != nullptr && kv) {
                v = *reinterpret_cast<const float4*>(ap[i] + k0);
                if (pp[i] != nullptr) {
                    const float4 p = __ldg(reinterpret_cast<const float4*>(pp[i] + k0));
                    v.x += p.x; v.y += p.y; v.z += p.z; v.w += p.w;
                }
            }
            ra[i] = v;
            float4 w = make_float4(0.f, 0.f, 0.f, 0.f);
            if (wp[i] != nullptr && kv) w = __ldg(reinterpret_cast<const float4*>(wp[i] + k0));
            rb[i] = w;
        }
    };
    auto sstore = [&](int buf) {
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            As[buf][kq4 + 0][srow[i]] = ra[i].x; As[buf][kq4 + 1][srow[i]] = ra[i].y;
            As[buf][kq4 + 2][srow[i]] = ra[i].z; As[buf][kq4 + 3][srow[i]] = ra[i].w;
            Bs[buf][kq4 + 0][srow[i]] = rb[i].x; Bs[buf][kq4 + 1][srow[i]] = rb[i].y;
            Bs[buf][kq4 + 2][srow[i]] = rb[i].z; Bs[buf][kq4 + 3][srow[i]] = rb[i].w;
        }
    };

    const int nk = (a.K + LBK - 1) / LBK;
    gload(0);
    sstore(0);
    __syncthreads();
    for (int kt = 0; kt < nk; ++kt) {
        const int buf = kt & 1;
        if (kt + 1 < nk) gload((kt + 1) * LBK);
#pragma unroll
        for (int kk = 0; kk < LBK; ++kk) {
            const float4 a0 = *reinterpret_cast<const float4*>(&As[buf][kk][ty * 4]);
            const float4 a1 = *reinterpret_cast<const float4*>(&As[buf][kk][64 + ty * 4]);
            const float4 b0 = *reinterpret_cast<const float4*>(&Bs[buf][kk][tx * 4]);
            const float4 b1 = *reinterpret_cast<const float4*>(&Bs[buf][kk][64 + tx * 4]);
            const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
            const float bv[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
        }
        if (kt + 1 < nk) {
            sstore(buf ^ 1);
            __syncthreads();
        }
    }

    // epilogue
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int lr = (i < 4) ? (ty * 4 + i) : (64 + ty * 4 + (i - 4));
        const int gr = m0 + lr;
        if (gr >= a.M) continue;
        const int orow = a.c_rows ? a.c_rows[gr] : gr;
#pragma unroll
        for (int half = 0; half < 2; ++half) {
            const int gn = n0 + half * 64 + tx * 4;
            if (gn >= a.N) continue;           // N is a multiple of 4
            float4 v = make_float4(acc[i][half * 4 + 0], acc[i][half * 4 + 1],
                                   acc[i][half * 4 + 2], acc[i][half * 4 + 3]);
            if (a.bias) {
                const float4 b = __ldg(reinterpret_cast<const float4*>(a.bias + gn));
                v.x += b.x; v.y += b.y; v.z += b.z; v.w += b.w;
            }
            if (a.relu) {
                v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f);
            }
            if (a.R) {
                const float4 r = *reinterpret_cast<const float4*>(a.R + (size_t)orow * a.ldr + gn);
                v.x = r.x + v.x; v.y = r.y + v.y; v.z = r.z + v.z; v.w = r.w + v.w;
            }
            *reinterpret_cast<float4*>(a.C + (size_t)orow * a.ldc + gn) = v;
        }
    }
}

// ------------------------------------------------------------------------------------------------
// layernorm: y[r] = (x[r] - mean) * rstd * gamma + beta, one warp per row, E multiple of 128, E <= 1024
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) layernorm_kernel(const float* __restrict__ x, const float* __restrict__ gamma,
                                                        const float* __restrict__ beta, float* __restrict__ y,
                                                        int M, int E, const int* stop) {
    FFB_PDL_SYNC();
    FFB_STOP_CHECK(stop);
    const int lane = threadIdx.x & 31;
    const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= M) return;
    const float* xr = x + (size_t)row * E;
    float4 v[8];
    const int nv = E >> 7;                    // float4 per lane
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        if (i < nv) {
            v[i] = *reinterpret_cast<const float4*>(xr + i * 128 + lane * 4);
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
    const float var = warp_sum(q) / (float)E;
    const float rstd = 1.0f / sqrtf(var + 1e-5f);
    float* yr = y + (size_t)row * E;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        if (i < nv) {
            const int c = i * 128 + lane * 4;
            const float4 g = __ldg(reinterpret_cast<const float4*>(gamma + c));
            const float4 b = __ldg(reinterpret_cast<const float4*>(beta + c));
            float4 o;
            o.x = v[i].x * rstd * g.x + b.x; o.y = v[i].y * rstd * g.y + b.y;
            o.z = v[i].z * rstd * g.z + b.z; o.w = v[i].w * rstd * g.w + b.w;
            *reinterpret_cast<float4*>(yr + c) = o;
        }
    }
}

// layernorm + operand formatting for the tensor-core GEMM: writes bf16x3 splits of y = LN(x) (out_plain) and/or of
// y + pos[r % pos_mod] (out_pos) -- with_pos_embed (transformer.py:144-145,205-206) fused into the producer.
template <int NV>      // float4 chunks per lane: E = 128 * NV columns (a compile-time bound keeps the row in NV * 4 registers: more CTAs per SM)
__global__ void __launch_bounds__(256) layernorm_split_kernel(const float* __restrict__ x, const float* __restrict__ gamma,
                                                              const float* __restrict__ beta, uint16_t* __restrict__ out_plain,
                                                              uint16_t* __restrict__ out_pos, long long split_stride,
                                                              const float* __restrict__ pos, int pos_mod,
                                                              int M, int E, int fmt, int* ovf, const int* stop,
                                                              const int* __restrict__ pos_idx = nullptr) {
    FFB_PDL_SYNC();
    FFB_STOP_CHECK(stop);
    const int lane = threadIdx.x & 31;
    const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= M) return;
    const float* xr = x + (size_t)row * E;
    float4 v[NV];
    const int nv = E >> 7;                      // <= NV
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        if (i < nv) {
            v[i] = *reinterpret_cast<const float4*>(xr + i * 128 + lane * 4);
            s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
        }
    }
    const float mean = warp_sum(s) / (float)E;
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        if (i < nv) {
            v[i].x -= mean; v[i].y -= mean; v[i].z -= mean; v[i].w -= mean;
            q += (v[i].x * v[i].x + v[i].y * v[i].y) + (v[i].z * v[i].z + v[i].w * v[i].w);
        }
    }
    const float var = warp_sum(q) / (float)E;
    const float rstd = 1.0f / sqrtf(var + 1e-5f);
    const float* prow = (out_pos != nullptr) ? pos + (size_t)(pos_idx ? pos_idx[row] : row % pos_mod) * E : nullptr;   // pos_idx: encoder rows
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        if (i < nv) {
            const int c = i * 128 + lane * 4;
            const float4 g = __ldg(reinterpret_cast<const float4*>(gamma + c));
            const float4 b = __ldg(reinterpret_cast<const float4*>(beta + c));
            float4 o;
            o.x = v[i].x * rstd * g.x + b.x; o.y = v[i].y * rstd * g.y + b.y;
            o.z = v[i].z * rstd * g.z + b.z; o.w = v[i].w * rstd * g.w + b.w;
            if (out_plain) store_split4(out_plain + (size_t)row * E + c, split_stride, o, fmt, ovf);
            if (out_pos) {
                const float4 pp = __ldg(reinterpret_cast<const float4*>(prow + c));
                o.x += pp.x; o.y += pp.y; o.z += pp.z; o.w += pp.w;
                store_split4(out_pos + (size_t)row * E + c, split_stride, o, fmt, ovf);
            }
        }
    }
}

// operand formatting without LayerNorm: out_plain <- split(x), out_pos <- split(x + pos[pos_idx[row]])  (A operands of the
// once-per-wireframe cross-attention K / V projections, transformer.py:248-251)
__global__ void __launch_bounds__(256) split_pos_kernel(const float* __restrict__ x, uint16_t* __restrict__ out_plain,
                                                        uint16_t* __restrict__ out_pos, long long split_stride,
                                                        const float* __restrict__ pos, const int* __restrict__ pos_idx,
                                                        int M, int E, int fmt, int* ovf) {
    const long long n4 = (long long)M * (E / 4);
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
        const int row = (int)(i / (E / 4)), c = (int)(i % (E / 4)) * 4;
        float4 v = *reinterpret_cast<const float4*>(x + (size_t)row * E + c);
        store_split4(out_plain + (size_t)row * E + c, split_stride, v, fmt, ovf);
        const float4 pp = __ldg(reinterpret_cast<const float4*>(pos + (size_t)pos_idx[row] * E + c));
        v.x += pp.x; v.y += pp.y; v.z += pp.z; v.w += pp.w;
        store_split4(out_pos + (size_t)row * E + c, split_stride, v, fmt, ovf);
    }
}

// operand formatting of a plain fp32 [M, E] array: out <- split(x) with the splits `split_stride` elements apart
__global__ void __launch_bounds__(256) split_rows_kernel(const float* __restrict__ x, uint16_t* __restrict__ out, long long split_stride,
                                                         int M, int E, int fmt, int* ovf) {
    const long long n4 = (long long)M * (E / 4);
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
        const float4 v = *reinterpret_cast<const float4*>(x + i * 4);
        store_split4(out + i * 4, split_stride, v, fmt, ovf);
    }
}

// ------------------------------------------------------------------------------------------------
// featurize: edge polylines -> [N, num_lines, P, 2] float32 + padding mask (datasets/data_para.py:8-25,59-68).
//   2-point edge: P points on the segment, x = x1 + (x2 - x1) * t with t = linspace(0, 1, P), evaluated in float64 exactly as
//   numpy does (multiply, then add: no FMA; t_i = i * (1 / (P - 1)), last = 1) and rounded to float32 on assignment;
//   longer polyline: points at indices linspace(0, n - 1, P).round(0) (round-half-even), cast to float32.
// One thread per (edge slot, sample); padded slots are zero-filled, mask = 1.
// ------------------------------------------------------------------------------------------------
__global__ void featurize_kernel(const double* __restrict__ pts, const long long* __restrict__ edge_off,
                                 const long long* __restrict__ wf_off, float* __restrict__ out, uint8_t* __restrict__ mask,
                                 long long* __restrict__ num_input, int N, int num_lines, int P) {
    const long long total = (long long)N * num_lines * P;
    for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
        const int i = (int)(idx % P);
        const long long slot = idx / P;
        const int e = (int)(slot % num_lines), w = (int)(slot / num_lines);
        const long long e0 = wf_off[w], ne = wf_off[w + 1] - e0;
        float2 o = make_float2(0.f, 0.f);
        if (e < ne) {
            const long long p0 = edge_off[e0 + e], n = edge_off[e0 + e + 1] - p0;
            if (n == 2) {
                const double t = (i == P - 1) ? 1.0 : __dmul_rn((double)i, 1.0 / (double)(P - 1));
                const double x1 = pts[2 * p0], y1 = pts[2 * p0 + 1], x2 = pts[2 * p0 + 2], y2 = pts[2 * p0 + 3];
                o.x = (float)__dadd_rn(x1, __dmul_rn(__dsub_rn(x2, x1), t));
                o.y = (float)__dadd_rn(y1, __dmul_rn(__dsub_rn(y2, y1), t));
            } else {
                const double step = (double)(n - 1) / (double)(P - 1);
                const double y = (i == P - 1) ? (double)(n - 1) : __dmul_rn((double)i, step);
                const long long j = (long long)rint(y);
                o.x = (float)pts[2 * (p0 + j)]; o.y = (float)pts[2 * (p0 + j) + 1];
            }
        }
        reinterpret_cast<float2*>(out)[idx] = o;
        if (i == 0) {
            mask[slot] = (e < ne) ? 0 : 1;
            if (e == 0) num_input[w] = ne;
        }
    }
}

// ------------------------------------------------------------------------------------------------
// parse_faces: predicted co-edge sequences -> canonical face loops, the step right after the path.
// One thread per predicted sequence (wireframe w, slot f):
//   1. Trainer.parse_parallel_faces, predict half (trainer.py:196-206): cut after the first token in [face_type_offset, token.len)
//      (the whole row if there is none), face type = that token - face_type_offset, subtract token.len, keep 0 <= v < num_edges;
//      an empty result is no face.
//   2. is_face_enclosed (dataset/tests/check_faces_enclosed.py:11-46): walk the edges; every edge must start where the previous one
//      ended (|dx| < tol and |dy| < tol on the polylines' first / last points, doubles), an edge that ends where the current loop
//      started closes the loop; the face is kept only if the last loop is closed.
//   3. filter_faces_by_encloseness (post_processing.py:8-20): every loop rolled so that its smallest index comes first (first
//      occurrence, like np.argmin), loops ordered by their first index (stable).
// Outputs per sequence: valid, face_type, n_loops, loop_len[<= T], indices[<= T] (loops concatenated in canonical order), n_indices.
// With check_enclosed = 0 only step 1 runs (n_loops = 0, indices in predicted order).
// ------------------------------------------------------------------------------------------------
constexpr int PF_MAX_T = 320;
__global__ void parse_faces_kernel(const long long* __restrict__ predict, int N, int F, int T, const double* __restrict__ pts,
                                   const long long* __restrict__ edge_off, const long long* __restrict__ wf_off, double tol,
                                   int check_enclosed, int num_token, int type_offset, uint8_t* __restrict__ valid,
                                   int* __restrict__ face_type, int* __restrict__ n_loops, int* __restrict__ loop_len,
                                   int* __restrict__ indices, int* __restrict__ n_indices) {
    const long long s = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (s >= (long long)N * F) return;
    const int w = (int)(s / F);
    const long long* p = predict + s * T;
    int* idx = indices + s * T;
    int* ll = loop_len + s * T;
    const long long e0 = wf_off[w], ne = wf_off[w + 1] - e0;
    int cut = T - 1;
    for (int t = 0; t < T; ++t) { const long long v = p[t]; if (v >= type_offset && v < num_token) { cut = t; break; } }
    const int ftype = (int)(p[cut] - type_offset);
    int n = 0;
    for (int t = 0; t <= cut; ++t) { const long long v = p[t] - num_token; if (v >= 0 && v < ne) idx[n++] = (int)v; }
    bool ok = n > 0;
    int nl = 0;
    if (ok && check_enclosed) {
        double cx = 0, cy = 0, lx = 0, ly = 0;            // start point of the open loop, end point of the previous edge
        bool open = false; int cur = 0;
        for (int i = 0; i < n && ok; ++i) {
            const long long a = edge_off[e0 + idx[i]], b = edge_off[e0 + idx[i] + 1] - 1;
            const double fx = pts[2 * a], fy = pts[2 * a + 1], ex = pts[2 * b], ey = pts[2 * b + 1];
            if (!open) { cx = fx; cy = fy; open = true; }
            else if (!(fabs(lx - fx) < tol && fabs(ly - fy) < tol)) { ok = false; break; }
            lx = ex; ly = ey; ++cur;
            if (fabs(ex - cx) < tol && fabs(ey - cy) < tol) { open = false; ll[nl++] = cur; cur = 0; }
        }
        if (open) ok = false;
        if (ok) {
            // roll every loop so that its smallest index comes first (in place, by cyclic rotation through a small stack buffer)
            int pos = 0;
            for (int l = 0; l < nl; ++l) {
                const int len = ll[l];
                int am = 0;
                for (int i = 1; i < len; ++i) if (idx[pos + i] < idx[pos + am]) am = i;
                if (am) {                                  // rotate left by am: three reversals
                    auto rev = [&](int a, int b) { while (a < b) { const int t2 = idx[pos + a]; idx[pos + a] = idx[pos + b]; idx[pos + b] = t2; ++a; --b; } };
                    rev(0, am - 1); rev(am, len - 1); rev(0, len - 1);
                }
                pos += len;
            }
            // order the loops by their first index: stable insertion sort of (start, len) with a scratch copy of the indices
            int tmp[PF_MAX_T]; int start[PF_MAX_T];
            pos = 0;
            for (int l = 0; l < nl; ++l) { start[l] = pos; pos += ll[l]; }
            for (int i = 0; i < n; ++i) tmp[i] = idx[i];
            for (int a = 1; a < nl; ++a) {
                const int st = start[a], le = ll[a], key = tmp[st];
                int b = a - 1;
                while (b >= 0 && tmp[start[b]] > key) { start[b + 1] = start[b]; ll[b + 1] = ll[b]; --b; }
                start[b + 1] = st; ll[b + 1] = le;
            }
            pos = 0;
            for (int l = 0; l < nl; ++l) for (int i = 0; i < ll[l]; ++i) idx[pos++] = tmp[start[l] + i];
        }
    }
    valid[s] = ok ? 1 : 0;
    face_type[s] = ftype;
    n_loops[s] = ok ? nl : 0;
    n_indices[s] = ok ? n : 0;
}

// ------------------------------------------------------------------------------------------------
// attention group geometry
// ------------------------------------------------------------------------------------------------
struct AttnGroups {
    int ragged;                 // 0: uniform groups, 1: per-group arrays
    // uniform: group g has queries [g*q_stride + q_off, +nq), keys [g*k_stride, +nk), outputs [g*o_stride, +nq)
    int nq, nk, q_stride, q_off, k_stride, o_stride;
    // ragged: queries (== output rows) [q_begin[g]*q_mul, q_begin[g+1]*q_mul); keys [k_begin[g], +k_len[g])
    const int* q_begin; int q_mul; const int* k_begin; const int* k_len;
    // split output (operand of the tensor-core out-projection): format and overflow flag
    int split_fmt; int* overflow;
    // teacher-forced pass (model_para.py:120,158-159; attn_rows_kernel only): causal mask (key position <= query position within the group)
    // and tgt_key_padding_mask (one byte per key row, non-zero = masked)
    int causal; const unsigned char* key_mask;
};

__device__ __forceinline__ void attn_group(const AttnGroups& g, int grp, long long& q0, int& nq,
                                           long long& k0, int& nk, long long& o0) {
    if (g.ragged) {
        q0 = (long long)g.q_begin[grp] * g.q_mul;
        nq = (int)((long long)g.q_begin[grp + 1] * g.q_mul - q0);
        k0 = g.k_begin[grp]; nk = g.k_len[grp]; o0 = q0;
    } else {
        q0 = (long long)grp * g.q_stride + g.q_off; nq = g.nq;
        k0 = (long long)grp * g.k_stride; nk = g.nk; o0 = (long long)grp * g.o_stride;
    }
}

// ------------------------------------------------------------------------------------------------
// attn_rows: warp-per-query-row attention for short key sets (decoder self-attention, P <= T).
// grid (group, head, q-tile of 64 rows); 4 warps; warp w owns rows w, w+4, ..., state in registers.
// ------------------------------------------------------------------------------------------------
constexpr int AR_BQ = 64, AR_BK = 32, AR_RPW = 16, AR_KS = 68;

__global__ void __launch_bounds__(128) attn_rows_kernel(const float* __restrict__ Q, int ldq,
                                                        const float* __restrict__ K, const float* __restrict__ V, int ldk,
                                                        float* __restrict__ O, int ldo, uint16_t* __restrict__ Os,
                                                        long long os_stride, const AttnGroups g, const int* stop) {
    FFB_STOP_CHECK(stop);
    __shared__ __align__(16) float Qs[AR_BQ][64];
    __shared__ __align__(16) float Ks[AR_BK][AR_KS];
    __shared__ __align__(16) float Vs[AR_BK][64];
    __shared__ __align__(16) float Ps[4][AR_BK];

    long long q0, k0, o0; int nq, nk;
    attn_group(g, blockIdx.x, q0, nq, k0, nk, o0);
    const int head = blockIdx.y;
    const int qt0 = blockIdx.z * AR_BQ;
    if (qt0 >= nq) return;
    const int nqt = min(AR_BQ, nq - qt0);
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;

    // stage Q tile (rows beyond nqt are zero)
    for (int idx = tid; idx < AR_BQ * 16; idx += 128) {
        const int r = idx >> 4, d4 = idx & 15;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (r < nqt) v = *reinterpret_cast<const float4*>(Q + (size_t)(q0 + qt0 + r) * ldq + head * 64 + d4 * 4);
        *reinterpret_cast<float4*>(&Qs[r][d4 * 4]) = v;
    }

    float m[AR_RPW], l[AR_RPW], oa[AR_RPW], ob[AR_RPW];
#pragma unroll
    for (int i = 0; i < AR_RPW; ++i) { m[i] = -INFINITY; l[i] = 0.f; oa[i] = 0.f; ob[i] = 0.f; }

    for (int kt = 0; kt < nk; kt += AR_BK) {
        __syncthreads();
        for (int idx = tid; idx < AR_BK * 16; idx += 128) {
            const int r = idx >> 4, d4 = idx & 15;
            float4 kv = make_float4(0.f, 0.f, 0.f, 0.f), vv = kv;
            if (kt + r < nk) {
                const size_t off = (size_t)(k0 + kt + r) * ldk + head * 64 + d4 * 4;
                kv = *reinterpret_cast<const float4*>(K + off);
                vv = *reinterpret_cast<const float4*>(V + off);
            }
            *reinterpret_cast<float4*>(&Ks[r][d4 * 4]) = kv;
            *reinterpret_cast<float4*>(&Vs[r][d4 * 4]) = vv;
        }
        __syncthreads();
        const bool kin = (kt + lane) < nk && !(g.key_mask != nullptr && g.key_mask[k0 + kt + lane] != 0);
#pragma unroll
        for (int i = 0; i < AR_RPW; ++i) {
            const int r = w + 4 * i;
            if (r < nqt) {                                   // warp-uniform
                const bool kvalid = kin && !(g.causal && (kt + lane) > (qt0 + r));      // generate_square_subsequent_mask (model_para.py:72-74)
                float s = 0.f;
#pragma unroll
                for (int d4 = 0; d4 < 16; ++d4) {
                    const float4 qv = *reinterpret_cast<const float4*>(&Qs[r][d4 * 4]);
                    const float4 kv = *reinterpret_cast<const float4*>(&Ks[lane][d4 * 4]);
                    s = fmaf(qv.x, kv.x, s); s = fmaf(qv.y, kv.y, s);
                    s = fmaf(qv.z, kv.z, s); s = fmaf(qv.w, kv.w, s);
                }
                s = kvalid ? s * 0.125f : -INFINITY;          // q * sqrt(1/64): exact power of two
                const float mn = fmaxf(m[i], warp_max(s));   // finite from the first tile on (key 0 is never masked: the anchor / SOS position)
                const float p = expf(s - mn);                 // masked lanes: exp(-inf) = 0
                const float corr = expf(m[i] - mn);           // first tile: exp(-inf) = 0
                l[i] = l[i] * corr + warp_sum(p);
                m[i] = mn;
                __syncwarp();
                Ps[w][lane] = p;
                __syncwarp();
                float a0 = oa[i] * corr, a1 = ob[i] * corr;
#pragma unroll
                for (int j4 = 0; j4 < AR_BK / 4; ++j4) {
                    const float4 pv = *reinterpret_cast<const float4*>(&Ps[w][j4 * 4]);
                    a0 = fmaf(pv.x, Vs[j4 * 4 + 0][lane], a0); a1 = fmaf(pv.x, Vs[j4 * 4 + 0][lane + 32], a1);
                    a0 = fmaf(pv.y, Vs[j4 * 4 + 1][lane], a0); a1 = fmaf(pv.y, Vs[j4 * 4 + 1][lane + 32], a1);
                    a0 = fmaf(pv.z, Vs[j4 * 4 + 2][lane], a0); a1 = fmaf(pv.z, Vs[j4 * 4 + 2][lane + 32], a1);
                    a0 = fmaf(pv.w, Vs[j4 * 4 + 3][lane], a0); a1 = fmaf(pv.w, Vs[j4 * 4 + 3][lane + 32], a1);
                }
                oa[i] = a0; ob[i] = a1;
            }
        }
    }
#pragma unroll
    for (int i = 0; i < AR_RPW; ++i) {
        const int r = w + 4 * i;
        if (r < nqt) {
            const float r0 = oa[i] / l[i], r1 = ob[i] / l[i];
            const size_t off = (size_t)(o0 + qt0 + r) * ldo + head * 64;
            if (Os == nullptr) {
                O[off + lane] = r0;
                O[off + lane + 32] = r1;
            } else {                                   // operand of the tensor-core out-projection
                store_split1(Os + off + lane, os_stride, r0, g.split_fmt, g.overflow);
                store_split1(Os + off + lane + 32, os_stride, r1, g.split_fmt, g.overflow);
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------
// attn_tiled: register-tiled flash attention (fp32), 64 queries x 64 keys per iteration, head dim 64.
// grid (q-tile, head, group); 128 threads; thread (ty = tid/8, tx = tid%8) owns rows ty+16i, keys tx+8j.
// ------------------------------------------------------------------------------------------------
constexpr int AT_BQ = 64, AT_BK = 64, AT_S = 68;
constexpr int AT_SMEM_BYTES = (3 * AT_BQ * AT_S + AT_BK * 64) * (int)sizeof(float);

__global__ void __launch_bounds__(128) attn_tiled_kernel(const float* __restrict__ Q, int ldq,
                                                         const float* __restrict__ K, const float* __restrict__ V, int ldk,
                                                         float* __restrict__ O, int ldo, uint16_t* __restrict__ Os,
                                                         long long os_stride, const AttnGroups g, const int* stop) {
    FFB_STOP_CHECK(stop);
    extern __shared__ __align__(16) float smem[];
    float (*Qs)[AT_S] = reinterpret_cast<float (*)[AT_S]>(smem);
    float (*Ks)[AT_S] = reinterpret_cast<float (*)[AT_S]>(smem + AT_BQ * AT_S);
    float (*Ps)[AT_S] = reinterpret_cast<float (*)[AT_S]>(smem + 2 * AT_BQ * AT_S);
    float (*Vs)[64] = reinterpret_cast<float (*)[64]>(smem + 3 * AT_BQ * AT_S);

    long long q0, k0, o0; int nq, nk;
    attn_group(g, blockIdx.z, q0, nq, k0, nk, o0);
    const int head = blockIdx.y;
    const int qt0 = blockIdx.x * AT_BQ;
    if (qt0 >= nq) return;
    const int nqt = min(AT_BQ, nq - qt0);
    const int tid = threadIdx.x, tx = tid & 7, ty = tid >> 3;

    for (int idx = tid; idx < AT_BQ * 16; idx += 128) {
        const int r = idx >> 4, d4 = idx & 15;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (r < nqt) v = *reinterpret_cast<const float4*>(Q + (size_t)(q0 + qt0 + r) * ldq + head * 64 + d4 * 4);
        *reinterpret_cast<float4*>(&Qs[r][d4 * 4]) = v;
    }

    float m[4], l[4], o[4][8];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        m[i] = -INFINITY; l[i] = 0.f;
#pragma unroll
        for (int j = 0; j < 8; ++j) o[i][j] = 0.f;
    }

    for (int kt = 0; kt < nk; kt += AT_BK) {
        __syncthreads();                                     // previous tile fully consumed (and Qs visible)
        for (int idx = tid; idx < AT_BK * 16; idx += 128) {
            const int r = idx >> 4, d4 = idx & 15;
            float4 kv = make_float4(0.f, 0.f, 0.f, 0.f), vv = kv;
            if (kt + r < nk) {
                const size_t off = (size_t)(k0 + kt + r) * ldk + head * 64 + d4 * 4;
                kv = *reinterpret_cast<const float4*>(K + off);
                vv = *reinterpret_cast<const float4*>(V + off);
            }
            *reinterpret_cast<float4*>(&Ks[r][d4 * 4]) = kv;
            *reinterpret_cast<float4*>(&Vs[r][d4 * 4]) = vv;
        }
        __syncthreads();

        // S = Q K^T for rows ty+16i, keys tx+8j
        float s[4][8];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 8; ++j) s[i][j] = 0.f;
#pragma unroll 4
        for (int d4 = 0; d4 < 16; ++d4) {
            float4 qv[4], kv[8];
#pragma unroll
            for (int i = 0; i < 4; ++i) qv[i] = *reinterpret_cast<const float4*>(&Qs[ty + 16 * i][d4 * 4]);
#pragma unroll
            for (int j = 0; j < 8; ++j) kv[j] = *reinterpret_cast<const float4*>(&Ks[tx + 8 * j][d4 * 4]);
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    s[i][j] = fmaf(qv[i].x, kv[j].x, s[i][j]); s[i][j] = fmaf(qv[i].y, kv[j].y, s[i][j]);
                    s[i][j] = fmaf(qv[i].z, kv[j].z, s[i][j]); s[i][j] = fmaf(qv[i].w, kv[j].w, s[i][j]);
                }
        }
        // online softmax
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            float mx = -INFINITY;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                s[i][j] = (kt + tx + 8 * j < nk) ? s[i][j] * 0.125f : -INFINITY;
                mx = fmaxf(mx, s[i][j]);
            }
            mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 1));
            mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 2));
            mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 4));
            const float mn = fmaxf(m[i], mx);                // finite: every tile has >= 1 valid key
            const float corr = expf(m[i] - mn);
            float ps = 0.f;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const float p = expf(s[i][j] - mn);
                ps += p;
                Ps[ty + 16 * i][tx + 8 * j] = p;
            }
            ps += __shfl_xor_sync(0xffffffffu, ps, 1);
            ps += __shfl_xor_sync(0xffffffffu, ps, 2);
            ps += __shfl_xor_sync(0xffffffffu, ps, 4);
            l[i] = l[i] * corr + ps;
            m[i] = mn;
#pragma unroll
            for (int j = 0; j < 8; ++j) o[i][j] *= corr;
        }
        __syncthreads();
        // O += P V for rows ty+16i, dims tx*8 .. tx*8+7
#pragma unroll 4
        for (int j4 = 0; j4 < AT_BK / 4; ++j4) {
            float4 pv[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) pv[i] = *reinterpret_cast<const float4*>(&Ps[ty + 16 * i][j4 * 4]);
#pragma unroll
            for (int jj = 0; jj < 4; ++jj) {
                const float4 v0 = *reinterpret_cast<const float4*>(&Vs[j4 * 4 + jj][tx * 8]);
                const float4 v1 = *reinterpret_cast<const float4*>(&Vs[j4 * 4 + jj][tx * 8 + 4]);
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const float p = (jj == 0) ? pv[i].x : (jj == 1) ? pv[i].y : (jj == 2) ? pv[i].z : pv[i].w;
                    o[i][0] = fmaf(p, v0.x, o[i][0]); o[i][1] = fmaf(p, v0.y, o[i][1]);
                    o[i][2] = fmaf(p, v0.z, o[i][2]); o[i][3] = fmaf(p, v0.w, o[i][3]);
                    o[i][4] = fmaf(p, v1.x, o[i][4]); o[i][5] = fmaf(p, v1.y, o[i][5]);
                    o[i][6] = fmaf(p, v1.z, o[i][6]); o[i][7] = fmaf(p, v1.w, o[i][7]);
                }
            }
        }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int r = ty + 16 * i;
        if (r < nqt) {
            const float inv = 1.0f / l[i];
            const size_t off = (size_t)(o0 + qt0 + r) * ldo + head * 64 + tx * 8;
            const float4 lo4 = make_float4(o[i][0] * inv, o[i][1] * inv, o[i][2] * inv, o[i][3] * inv);
            const float4 hi4 = make_float4(o[i][4] * inv, o[i][5] * inv, o[i][6] * inv, o[i][7] * inv);
            if (Os == nullptr) {
                *reinterpret_cast<float4*>(O + off) = lo4;
                *reinterpret_cast<float4*>(O + off + 4) = hi4;
            } else {
                store_split4(Os + off, os_stride, lo4, g.split_fmt, g.overflow);
                store_split4(Os + off + 4, os_stride, hi4, g.split_fmt, g.overflow);
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------
// small data-movement kernels
// ------------------------------------------------------------------------------------------------
// x[b*P + p] = mem[row_off[seq_wf[b]] + tok[p*B + b]]   (model_para.py:217-219)
__global__ void gather_tgt_kernel(const float* __restrict__ mem, const int* __restrict__ row_off,
                                  const int* __restrict__ seq_wf, const int* __restrict__ tok,
                                  float* __restrict__ x, int B, int P, int E, const int* stop) {
    FFB_PDL_SYNC();
    FFB_STOP_CHECK(stop);
    const int e4n = E >> 2;
    const long long total = (long long)B * P * e4n;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int c = (int)(i % e4n);
        const long long r = i / e4n;
        const int b = (int)(r / P), p = (int)(r % P);
        const int src = row_off[seq_wf[b]] + tok[(size_t)p * B + b];
        reinterpret_cast<float4*>(x)[r * e4n + c] = reinterpret_cast<const float4*>(mem)[(size_t)src * e4n + c];
    }
}

// dst[r] = src[r*mul + off]  (rows of E floats)
__global__ void copy_rows_kernel(const float* __restrict__ src, float* __restrict__ dst, int rows, int mul, int off,
                                 int E, const int* stop) {
    FFB_PDL_SYNC();
    FFB_STOP_CHECK(stop);
    const int e4n = E >> 2;
    const long long total = (long long)rows * e4n;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int c = (int)(i % e4n);
        const long long r = i / e4n;
        reinterpret_cast<float4*>(dst)[r * e4n + c] = reinterpret_cast<const float4*>(src)[((size_t)r * mul + off) * e4n + c];
    }
}

// rows 0..3 of every wireframe's packed block <- special-token table (embedding.py:30-32,36)
__global__ void token_rows_kernel(const float* __restrict__ table, const int* __restrict__ row_off, float* __restrict__ x,
                                  int n_wf, int num_token, int E) {
    const int e4n = E >> 2;
    const int total = n_wf * num_token * e4n;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        const int c = i % e4n;
        const int t = (i / e4n) % num_token;
        const int w = i / (e4n * num_token);
        reinterpret_cast<float4*>(x)[(size_t)(row_off[w] + t) * e4n + c] = reinterpret_cast<const float4*>(table)[t * e4n + c];
    }
}

// tok[0*B + b] = first token of sequence b (anchor or SOS); also resets the loop state
__global__ void init_tokens_kernel(const int* __restrict__ seq_first, int* __restrict__ tok, int B,
                                   int* stop, int* steps_run, int* eos_found) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < B) tok[i] = seq_first[i];
    if (i == 0) { *stop = 0; *steps_run = 0; *eos_found = 0; }
}

// ------------------------------------------------------------------------------------------------
// pointer head: logits[b, j] = <mem[row_off[w]+j], ptr[b]>, j < v_len[w]; first-max argmax; append.
// One CTA (8 warps) per sequence; warp w scans rows w, w+8, ...; fixed reduction order.
// ------------------------------------------------------------------------------------------------
struct PointerArgs {
    const float* mem;            // packed memory [R,E]
    const float* ptr; int ptr_stride_rows; int ptr_off;   // pointer row of sequence b: ptr[(b*stride + off) * E]
    int ldm;                     // row stride of `mem` in floats (E, or E + 4 for the folded head)
    int bias_col;                // HD instantiations: logit += mem[row][bias_col] (the folded project bias), cross-lane reduction in float64
    const int* row_off; const int* v_len; const int* seq_wf;
    float* logits; int L;        // [B, L] masked logits (finfo.min beyond v_len), or null
    int* tok_out;                // tok[(P)*B + b] slot for the new token, or null (forced-prefix mode)
    int B, E;
    int num_token;
    int* nonstop_count;          // parallel: number of sequences with next >= num_token this step
    int* eos_count;              // seq2seq : cumulative count of next == EOS
    const int* stop;
};

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// HD ("folded head"): mem = memory . [W_project^T ; b_project] precomputed per wireframe in float64 (rounded to fp32),
// ptr = the float64 LayerNorm of the last position (rounded to fp32): logit = <mem[row, :E], ptr> + mem[row, E].  The 16-term per-lane
// chains stay fp32 FMAs on small partial sums; the cross-lane reduction runs in float64.
template <bool HD>
__global__ void __launch_bounds__(256) pointer_kernel(const PointerArgs a) {
    FFB_PDL_SYNC();
    FFB_STOP_CHECK(a.stop);
    __shared__ __align__(16) float ps[1024];
    __shared__ float bestv[8];
    __shared__ int besti[8];
    const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    const float* pr = a.ptr + ((size_t)b * a.ptr_stride_rows + a.ptr_off) * a.E;
    for (int c = tid; c < a.E; c += 256) ps[c] = pr[c];
    __syncthreads();
    const int wf = a.seq_wf[b];
    const int r0 = a.row_off[wf], vl = a.v_len[wf];
    float bv = -INFINITY; int bi = 0x7fffffff;
    for (int j = w; j < vl; j += 8) {
        const float* mr = a.mem + (size_t)(r0 + j) * a.ldm;
        float sd = 0.f;
        for (int c = lane * 4; c < a.E; c += 128) {
            const float4 mv = *reinterpret_cast<const float4*>(mr + c);
            const float4 pv = *reinterpret_cast<const float4*>(&ps[c]);
            sd = fmaf(mv.x, pv.x, sd); sd = fmaf(mv.y, pv.y, sd); sd = fmaf(mv.z, pv.z, sd); sd = fmaf(mv.w, pv.w, sd);
        }
        const float s = HD ? (float)(warp_sum((double)sd) + (double)mr[a.bias_col]) : warp_sum(sd);
        if (a.logits && lane == 0) a.logits[(size_t)b * a.L + j] = s;
        // strictly greater: first max within this warp's rows; like torch.argmax a NaN counts as the maximum (first NaN wins)
        if (bi == 0x7fffffff || s > bv || (s != s && bv == bv)) { bv = s; bi = j; }
    }
    if (a.logits) for (int j = vl + tid; j < a.L; j += 256) a.logits[(size_t)b * a.L + j] = -FLT_MAX;   // finfo.min
    if (lane == 0) { bestv[w] = bv; besti[w] = bi; }
    __syncthreads();
    if (tid == 0) {
        float v = bestv[0]; int idx = besti[0];
        for (int k = 1; k < 8; ++k) {
            if (besti[k] == 0x7fffffff) continue;                  // this warp had no rows
            const bool vn = (v != v), kn = (bestv[k] != bestv[k]);
            if ((kn && (!vn || besti[k] < idx)) || (!vn && !kn && (bestv[k] > v || (bestv[k] == v && besti[k] < idx)))) { v = bestv[k]; idx = besti[k]; }
        }
        if (a.tok_out) {
            a.tok_out[b] = idx;
            if (a.nonstop_count && idx >= a.num_token) atomicAdd(a.nonstop_count, 1);
            if (a.eos_count && idx == 3) atomicAdd(a.eos_count, 1);        // token.EOS == 3 (config.py:44)
        }
    }
}

// After each step: evaluate the reference's stop predicate and count the step.
//   parallel (model_para.py:232): stop if all next < num_token  <=> nonstop_count == 0
//   seq2seq  (model.py:207-210) : stop if cumulative EOS count == N
// Batched variant: the sequences of one wireframe share its memory rows, so a CTA takes (wireframe, PB_SEQ of its sequences) and streams the
// memory rows ONCE for all 16 (pointer_kernel re-reads them per sequence: ~1 GB through L2 per step at the bench size).  Per
// (row, sequence) the dot product is evaluated in exactly the order of pointer_kernel (per-lane fmaf chain, then warp_sum), so logits and
// tokens are bit-identical to it.  grid (ceil(max sequences per wireframe / PB_SEQ), N), 256 threads, dynamic smem PB_SEQ * E floats.
constexpr int PB_SEQ = 4;
template <bool HD>
__global__ void __launch_bounds__(256) pointer_batched_kernel(const PointerArgs a, const int* __restrict__ seq_off) {
    FFB_PDL_SYNC();
    FFB_STOP_CHECK(a.stop);
    extern __shared__ __align__(16) float pb_smem[];
    float* ps = pb_smem;                                     // [PB_SEQ][E]
    __shared__ float bestv[8][PB_SEQ];
    __shared__ int besti[8][PB_SEQ];
    const int wf = blockIdx.y, tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    const int s0 = seq_off[wf] + blockIdx.x * PB_SEQ;
    const int nq = min(PB_SEQ, seq_off[wf + 1] - s0);
    if (nq <= 0) return;
    for (int i = tid; i < nq * a.E; i += 256) {
        const int q = i / a.E, c = i - q * a.E;
        ps[q * a.E + c] = a.ptr[((size_t)(s0 + q) * a.ptr_stride_rows + a.ptr_off) * a.E + c];
    }
    __syncthreads();
    const int r0 = a.row_off[wf], vl = a.v_len[wf];
    const int nv = a.E >> 7;                                 // float4 chunks per lane (E multiple of 128, <= 1024)
    float bv = -INFINITY; int bi = 0x7fffffff;               // lane q < nq tracks the running best of sequence q over this warp's rows
    for (int j = w; j < vl; j += 8) {
        const float* mr = a.mem + (size_t)(r0 + j) * a.ldm;
        float4 mv[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) if (i < nv) mv[i] = *reinterpret_cast<const float4*>(mr + lane * 4 + i * 128);
        float sacc[PB_SEQ];
#pragma unroll
        for (int q = 0; q < PB_SEQ; ++q) {                   // independent fmaf chains (same per-chain order as pointer_kernel)
            sacc[q] = 0.f;
            if (q < nq) {
                const float* pq = ps + q * a.E + lane * 4;
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    if (i < nv) {
                        const float4 pv = *reinterpret_cast<const float4*>(pq + i * 128);
                        sacc[q] = fmaf(mv[i].x, pv.x, sacc[q]); sacc[q] = fmaf(mv[i].y, pv.y, sacc[q]);
                        sacc[q] = fmaf(mv[i].z, pv.z, sacc[q]); sacc[q] = fmaf(mv[i].w, pv.w, sacc[q]);
                    }
                }
            }
        }
        const float mbias = HD ? mr[a.bias_col] : 0.f;
#pragma unroll
        for (int q = 0; q < PB_SEQ; ++q) sacc[q] = HD ? (float)(warp_sum((double)sacc[q]) + (double)mbias) : warp_sum(sacc[q]);
#pragma unroll
        for (int q = 0; q < PB_SEQ; ++q) {
            if (q < nq && lane == q) {
                const float sv = sacc[q];
                if (a.logits) a.logits[(size_t)(s0 + q) * a.L + j] = sv;
                if (bi == 0x7fffffff || sv > bv || (sv != sv && bv == bv)) { bv = sv; bi = j; }      // first max; NaN counts as the maximum
            }
        }
    }
    if (a.logits)
        for (int q = 0; q < nq; ++q)
            for (int j = vl + tid; j < a.L; j += 256) a.logits[(size_t)(s0 + q) * a.L + j] = -FLT_MAX;   // finfo.min
    if (lane < PB_SEQ) { bestv[w][lane] = bv; besti[w][lane] = bi; }
    __syncthreads();
    if (tid < nq) {
        float v = bestv[0][tid]; int idx = besti[0][tid];
        for (int k = 1; k < 8; ++k) {
            if (besti[k][tid] == 0x7fffffff) continue;             // this warp had no rows
            const float kv = bestv[k][tid]; const int ki = besti[k][tid];
            const bool vn = (v != v), kn = (kv != kv);
            if (idx == 0x7fffffff || (kn && (!vn || ki < idx)) || (!vn && !kn && (kv > v || (kv == v && ki < idx)))) { v = kv; idx = ki; }
        }
        if (a.tok_out) {
            a.tok_out[s0 + tid] = idx;
            if (a.nonstop_count && idx >= a.num_token) atomicAdd(a.nonstop_count, 1);
            if (a.eos_count && idx == 3) atomicAdd(a.eos_count, 1);        // token.EOS == 3 (config.py:44)
        }
    }
}

// ------------------------------------------------------------------------------------------------
// beam search step (BASELINE.json configs[3]; specification: oracle/beam_oracle.py -- the reference has no beam search).
// One CTA per anchor, one warp per hypothesis w < W: float64 log-sum-exp over the un-masked logits, the W best rows by
// (logit desc, row asc); thread 0 merges the <= W*W candidates by (cum + logp desc, hypothesis asc, rank asc); then the token
// histories are re-ordered into the other token buffer and the new tokens appended.  W = 1 degenerates to first-max argmax.
// ------------------------------------------------------------------------------------------------
constexpr int BEAM_MAX = 8;
__global__ void beam_init_kernel(double* __restrict__ cum, int n, int W) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) cum[i] = (i % W == 0) ? 0.0 : -INFINITY;
}

__global__ void __launch_bounds__(32 * BEAM_MAX) beam_step_kernel(const float* __restrict__ logits, int L, const int* __restrict__ seq_wf,
                                                                  const int* __restrict__ v_len, double* __restrict__ cum,
                                                                  const int* __restrict__ tok_old, int* __restrict__ tok_new, int P, int T,
                                                                  int Btot, int W, int num_token, int* nonstop_count, const int* stop) {
    FFB_PDL_SYNC();
    const int a = blockIdx.x, tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    if (stop != nullptr && *stop != 0) {            // after an early stop both token buffers are kept identical (the host keeps flipping them)
        for (int i = tid; i < T * W; i += blockDim.x) { const int p = i / W, j = i % W; tok_new[(size_t)p * Btot + a * W + j] = tok_old[(size_t)p * Btot + a * W + j]; }
        return;
    }
    __shared__ double c_score[BEAM_MAX][BEAM_MAX];
    __shared__ int c_tok[BEAM_MAX][BEAM_MAX];
    __shared__ int c_n[BEAM_MAX];
    __shared__ int s_parent[BEAM_MAX], s_token[BEAM_MAX];
    __shared__ double s_cum[BEAM_MAX];
    const int sq = a * W + w;
    const double cumw = cum[sq];
    const int vl = v_len[seq_wf[sq]];
    const float* lg = logits + (size_t)sq * L;
    int n_c = 0;
    if (cumw > -INFINITY) {
        float m = -INFINITY;
        for (int j = lane; j < vl; j += 32) m = fmaxf(m, lg[j]);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
        double sum = 0.0;
        for (int j = lane; j < vl; j += 32) sum += exp((double)lg[j] - (double)m);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
        const double lse = (double)m + log(sum);
        int taken[BEAM_MAX];
        for (int r = 0; r < W && r < vl; ++r) {
            float bv = -INFINITY; int bi = 0x7fffffff;
            for (int j = lane; j < vl; j += 32) {
                bool skip = false;
                for (int q = 0; q < r; ++q) skip |= (taken[q] == j);
                const float v = lg[j];
                if (!skip && (bi == 0x7fffffff || v > bv)) { bv = v; bi = j; }       // lane-local: ascending j, strict > keeps the first
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
                const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
                if (oi != 0x7fffffff && (bi == 0x7fffffff || ov > bv || (ov == bv && oi < bi))) { bv = ov; bi = oi; }
            }
            taken[r] = bi;
            if (lane == 0) { c_score[w][r] = cumw + ((double)bv - lse); c_tok[w][r] = bi; }
            ++n_c;
        }
    }
    if (lane == 0) c_n[w] = n_c;
    __syncthreads();
    if (tid == 0) {
        bool used[BEAM_MAX][BEAM_MAX];
        for (int i = 0; i < W; ++i) for (int r = 0; r < W; ++r) used[i][r] = false;
        for (int j = 0; j < W; ++j) {
            int bw = -1, br = -1; double bs = 0.0;
            for (int i = 0; i < W; ++i)
                for (int r = 0; r < c_n[i]; ++r)
                    if (!used[i][r] && (bw < 0 || c_score[i][r] > bs)) { bw = i; br = r; bs = c_score[i][r]; }
            if (bw < 0) { s_parent[j] = 0; s_token[j] = 0; s_cum[j] = -INFINITY; continue; }     // fewer candidates than beams: dead hypothesis
            used[bw][br] = true;
            s_parent[j] = bw; s_token[j] = c_tok[bw][br]; s_cum[j] = bs;
        }
    }
    __syncthreads();
    for (int i = tid; i < P * W; i += blockDim.x) {
        const int p = i / W, j = i % W;
        tok_new[(size_t)p * Btot + a * W + j] = tok_old[(size_t)p * Btot + a * W + s_parent[j]];
    }
    if (tid < W) {
        tok_new[(size_t)P * Btot + a * W + tid] = s_token[tid];
        cum[a * W + tid] = s_cum[tid];
        if (nonstop_count && s_token[tid] >= num_token) atomicAdd(nonstop_count, 1);
    }
}

// beams [slot, w, t] (int64) and scores [slot, w] from the token buffer: slot_seq[slot] = hypothesis 0 of the anchor feeding the slot
__global__ void expand_beams_kernel(const int* __restrict__ tok, const double* __restrict__ cum, const int* __restrict__ slot_seq,
                                    const int* __restrict__ steps_run, long long* __restrict__ beams, double* __restrict__ scores,
                                    long long n_slots, int Btot, int W, int T) {
    const long long total = n_slots * W * T;
    const int filled = *steps_run + 1;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int t = (int)(i % T);
        const int w = (int)((i / T) % W);
        const long long s = i / ((long long)T * W);
        beams[i] = (t < filled) ? (long long)tok[(size_t)t * Btot + slot_seq[s] + w] : 0ll;
        if (t == 0) scores[s * W + w] = cum[slot_seq[s] + w];
    }
}

// Batch splitting over GPUs (BASELINE.json configs[2]): the reference's stop predicate `all(next < 4)` (model_para.py:232) ranges over
// the WHOLE batch, whose wireframes are now spread over `world` ranks.  Every rank publishes its local verdict for this step into the
// flag buffer of every rank (plain stores through peer mappings: NVLink P2P / CUDA IPC), waits until all verdicts of the step have
// arrived in its own buffer and combines them -- one word per rank and step, no host round trip, no NCCL call on the data path.
//   slot value = (epoch << 2) | (1 = some sequence of the rank continues, 2 = all of them emitted special tokens)
//   buffers are double-buffered by epoch parity (a rank can be at most one decode ahead of its slowest peer)
constexpr int XCHG_MAXW = 16;
struct XchgArgs { int* peers[XCHG_MAXW]; int rank, world, T; };
__global__ void step_end_xchg_kernel(int* nonstop_count, int* stop, int* steps_run, const XchgArgs x, int step, int epoch, int* err) {
    FFB_PDL_SYNC();
    if (*stop) return;
    const int lane = threadIdx.x;
    const int local = (*nonstop_count == 0) ? 2 : 1;
    const int slot = ((epoch & 1) * x.T + step) * XCHG_MAXW;
    if (lane < x.world) {
        volatile int* dst = x.peers[lane] + slot + x.rank;
        *dst = (epoch << 2) | local;
    }
    __threadfence_system();
    bool all = true, timed_out = false;
    if (lane < x.world) {
        volatile int* src = x.peers[x.rank] + slot + lane;
        unsigned long long t0;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
        int v = *src;
        while ((v >> 2) != epoch) {
            unsigned long long t1;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
            if (t1 - t0 > 20000000000ull) { timed_out = true; break; }      // 20 s: a peer died; fail instead of hanging the GPU
            __nanosleep(200);
            v = *src;
        }
        all = ((v & 3) == 2);
    }
    all = __all_sync(0xffffffffu, all);
    timed_out = __any_sync(0xffffffffu, timed_out);
    if (lane == 0) {
        *nonstop_count = 0;
        *steps_run += 1;
        if (timed_out) { *err = 1; *stop = 1; }
        else if (all) *stop = 1;
    }
}

__global__ void step_end_kernel(int mode, int n_seq, int* nonstop_count, int* eos_count, int* stop, int* steps_run) {
    FFB_PDL_SYNC();
    if (*stop) return;
    *steps_run += 1;
    if (mode == 0) {
        if (*nonstop_count == 0) *stop = 1;
        *nonstop_count = 0;
    } else {
        if (*eos_count == n_seq) *stop = 1;
    }
}

// predict[w, f, t] (int64) from the step-major token buffer; zero after the executed steps.
//   slot_seq[w*F + f] = decoded sequence feeding output slot (w,f)
__global__ void expand_predict_kernel(const int* __restrict__ tok, const int* __restrict__ slot_seq,
                                      const int* __restrict__ steps_run, long long* __restrict__ predict,
                                      long long n_slots, int B, int T) {
    const long long total = n_slots * T;
    const int filled = *steps_run + 1;             // rows of `predicts` that exist (start token + S steps)
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int t = (int)(i % T);
        const long long s = i / T;
        predict[i] = (t < filled) ? (long long)tok[(size_t)t * B + slot_seq[s]] : 0ll;
    }
}

// logits_out[slot, :] = logits[slot_seq[slot], :]
__global__ void expand_rows_kernel(const float* __restrict__ src, const int* __restrict__ slot_seq,
                                   float* __restrict__ dst, long long n_slots, int L) {
    const long long total = n_slots * L;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int c = (int)(i % L);
        const long long s = i / L;
        dst[i] = src[(size_t)slot_seq[s] * L + c];
    }
}

// padded [N, L, E] view of the packed memory (zeros in padded rows)
__global__ void unpack_memory_kernel(const float* __restrict__ mem, const int* __restrict__ row_off,
                                     const int* __restrict__ v_len, float* __restrict__ out, int n_wf, int L, int E) {
    const int e4n = E >> 2;
    const long long total = (long long)n_wf * L * e4n;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int c = (int)(i % e4n);
        const long long r = i / e4n;
        const int j = (int)(r % L), w = (int)(r / L);
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (j < v_len[w]) v = reinterpret_cast<const float4*>(mem)[(size_t)(row_off[w] + j) * e4n + c];
        reinterpret_cast<float4*>(out)[i] = v;
    }
}

// Per-row plan of the packed memory, expanded on the device from the wireframes' row offsets (the host only scans the masks):
//   pos_idx[r]  = row index within its wireframe (position-table row, embedding.py:106-108)
//   edge e (valid edges in wireframe order) : edge_src[e] = its slot in the [N, num_lines] input, edge_dst[e] = its memory row
__global__ void plan_rows_kernel(const int* __restrict__ row_off, int n_wf, int R, int num_lines, int num_token,
                                 int* __restrict__ pos_idx, int* __restrict__ edge_src, int* __restrict__ edge_dst) {
    for (int r = blockIdx.x * blockDim.x + threadIdx.x; r < R; r += gridDim.x * blockDim.x) {
        int lo = 0, hi = n_wf - 1;                       // largest w with row_off[w] <= r
        while (lo < hi) { const int mid = (lo + hi + 1) >> 1; if (row_off[mid] <= r) lo = mid; else hi = mid - 1; }
        const int j = r - row_off[lo];
        pos_idx[r] = j;
        if (j >= num_token) {
            const int e = r - num_token * (lo + 1);      // every wireframe before (and this one) contributes num_token non-edge rows
            edge_src[e] = lo * num_lines + (j - num_token);
            edge_dst[e] = r;
        }
    }
}

// Layer-0 q/k/v of the decoder do not depend on the step for positions that already exist (the layer-0 input of position p is
// memory[token_p] and its LayerNorm + in-projection involve that row alone; tokens never change once appended): they are computed
// for the NEW position only and kept in a position-stable cache [2][B][T][ld].  This kernel builds the (sequence, position)-ordered
// operand rows a_qkv[(b*P + p)] the attention kernel reads: old positions from the cache, the new one from `fresh` [2][capb][ld]
// (which it also appends to the cache).  16-byte chunks; `parts` are the fp16x2 halves.
__global__ void assemble_qkv0_kernel(const uint16_t* __restrict__ fresh, long long fresh_stride, uint16_t* __restrict__ cache, long long cache_stride,
                                     uint16_t* __restrict__ out, long long out_stride, int B, int P, int T, int ld, const int* stop) {
    FFB_PDL_SYNC();
    FFB_STOP_CHECK(stop);
    const int c8n = ld >> 3;
    const long long total = 2ll * B * P * c8n;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int c = (int)(i % c8n);
        long long r = i / c8n;
        const int p = (int)(r % P); r /= P;
        const int b = (int)(r % B); const int part = (int)(r / B);
        uint4 v;
        uint16_t* cslot = cache + (size_t)part * cache_stride + ((size_t)b * T + p) * ld + c * 8;
        if (p == P - 1) {
            v = *reinterpret_cast<const uint4*>(fresh + (size_t)part * fresh_stride + (size_t)b * ld + c * 8);
            *reinterpret_cast<uint4*>(cslot) = v;
        } else {
            v = *reinterpret_cast<const uint4*>(cslot);
        }
        *reinterpret_cast<uint4*>(out + (size_t)part * out_stride + ((size_t)b * P + p) * ld + c * 8) = v;
    }
}

// dst [rows, cols_out] = src [rows, cols_in] zero-padded on the right
__global__ void pad_cols_kernel(const float* __restrict__ src, float* __restrict__ dst, int rows, int cols_in, int cols_out) {
    const long long total = (long long)rows * cols_out;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int c = (int)(i % cols_out); const long long r = i / cols_out;
        dst[i] = c < cols_in ? src[r * cols_in + c] : 0.f;
    }
}

// coordinates of the valid edges -> fp16x2 operand rows of 128 columns (zero-padded beyond in_dim) at the edges' memory rows
__global__ void split_coords_kernel(const float* __restrict__ coords, const int* __restrict__ edge_src, const int* __restrict__ edge_dst,
                                    uint16_t* __restrict__ out, long long split_stride, int Re, int in_dim, int* ovf) {
    const long long total = (long long)Re * 32;          // 32 float4 chunks per row
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int c = (int)(i & 31) * 4; const long long e = i >> 5;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (c + 3 < in_dim) v = *reinterpret_cast<const float4*>(coords + (size_t)edge_src[e] * in_dim + c);
        else if (c < in_dim) {
            const float* p = coords + (size_t)edge_src[e] * in_dim + c;
            v.x = p[0]; if (c + 1 < in_dim) v.y = p[1]; if (c + 2 < in_dim) v.z = p[2];
        }
        store_split4(out + (size_t)edge_dst[e] * 128 + c, split_stride, v, 2, ovf);
    }
}

// dst[c][r] = src[r][c]  (src [rows, cols])
__global__ void transpose_kernel(const float* __restrict__ src, float* __restrict__ dst, int rows, int cols) {
    const long long total = (long long)rows * cols;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int c = (int)(i / rows), r = (int)(i % rows);
        dst[i] = src[(size_t)r * cols + c];
    }
}

// inverse of unpack_memory_kernel: [N, L, E] -> packed valid rows (parity hook ffb_set_memory)
__global__ void pack_memory_kernel(const float* __restrict__ in, const int* __restrict__ row_off, const int* __restrict__ v_len,
                                   float* __restrict__ mem, int n_wf, int L, int E) {
    const int e4n = E >> 2;
    for (int w = 0; w < n_wf; ++w) {
        const long long total = (long long)v_len[w] * e4n;
        for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
            const int c = (int)(i % e4n), j = (int)(i / e4n);
            reinterpret_cast<float4*>(mem)[(size_t)(row_off[w] + j) * e4n + c] = reinterpret_cast<const float4*>(in)[((size_t)w * L + j) * e4n + c];
        }
    }
}

// int64 prefix [P, B_full] -> int32 tok [P, B_eff] through seq -> first slot map
// teacher forcing: tok[p*B + b] = label[wf, f, p], kmask[b*P + p] = label_mask[wf, f, p] for the P = T - 1 decoder input positions of sequence
// b = (wireframe wf, anchor slot f) (model_para.py:84-86: tgt = label[..., :-1])
__global__ void load_labels_kernel(const long long* __restrict__ label, const unsigned char* __restrict__ label_mask, int label_rows, int T,
                                   const int* __restrict__ seq_wf, const int* __restrict__ seq_off, int* __restrict__ tok,
                                   unsigned char* __restrict__ kmask, int B, int P) {
    const long long total = (long long)B * P;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int b = (int)(i / P), p = (int)(i % P);
        const int wf = seq_wf[b], f = b - seq_off[wf];
        const size_t src = ((size_t)wf * label_rows + f) * T + p;
        tok[(size_t)p * B + b] = (int)label[src];
        kmask[i] = label_mask[src];
    }
}

__global__ void load_prefix_kernel(const long long* __restrict__ prefix, const int* __restrict__ seq_slot,
                                   int* __restrict__ tok, int P, int B_full, int B) {
    const int total = P * B;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        const int p = i / B, b = i % B;
        tok[i] = (int)prefix[(size_t)p * B_full + seq_slot[b]];
    }
}

}  // namespace ffb
