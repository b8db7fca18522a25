// attn_x.cuh -- decoder CROSS-attention core on the 5th-generation tensor cores (tcgen05 + TMEM + TMA), sm_100a only.
//
//   O = softmax(Q K^T / 8) V   per (wireframe, head): the queries are all prefix positions of all sequences of one
//   wireframe (transformer.py:247-251 with memory_key_padding_mask realised as "only valid rows exist"), the keys /
//   values are the wireframe's rows of the cross-attention cache (computed once per wireframe).
//
// Operands are fp16x2 splits (x = h + l), every product is  l*h + h*l + h*h  (3 MMAs, fp32 accumulate in TMEM):
//   Q   a_qc  [2][rows][E]            written by the query-projection GEMM epilogue       -> A of  S = Q K^T
//   K   kc_h  [2][R][Ld*E]            split once per wireframe at encode time             -> B of  S   (K-major: head dim)
//   Vt  vt_h  [2][Ld*E][Rp]           TRANSPOSED value cache, keys contiguous, every      -> B of  O = P V (K-major: keys)
//                                     wireframe's key range padded with zeros to 32 keys
//   P   softmax weights, written by the softmax warps into shared memory as fp16x2      -> A of  O
// All smem operand tiles are 64-byte rows with SWIZZLE_64B (the layout gemm_tc.cuh uses), 32 contraction elements per tile.
//
// One persistent CTA per SM walks a contiguous range of work items (wireframe, head, 128-query tile); K / Vt stay in
// shared memory while consecutive items share (wireframe, head).  Roles (192 threads):
//   warp 0      TMA producer (Q per item, K / Vt per (wireframe, head))
//   warp 1      TMEM allocator + MMA issuer (one thread): S = Q K^T (N = keys rounded to 16, <= 256), then O += P_c Vt_c per
//               32-key chunk as the softmax warps publish P_c
//   warps 2-5   softmax + epilogue, one query row per thread (TMEM lane = row): pass 1 row max over S, pass 2
//               p = 2^(s*log2e/8 - m + 12) per 32-key chunk -> fp16x2 -> smem (double-buffered), finally O / l -> fp16x2 ->
//               coalesced global stores through a per-warp staging area.
// Keys <= 256 per wireframe (ours.yml: 220); larger geometries use attn_h_kernel.
#pragma once
#include "gemm_tc.cuh"
#include "attn_h.cuh"   // split_pair, ex2_approx

namespace ffb {
namespace ax {

constexpr int BQ = 128, KMAX = 256, KC = 32, NUM_THREADS = 192, MAX_GROUPS = 1023;
constexpr int Q_TILE = BQ * 64;                  // bytes of one (part, k-chunk) Q tile: 128 rows x 64 B
constexpr int K_TILE = KMAX * 64;                // 256 rows x 64 B
constexpr int V_TILE = 64 * 64;                  // one 32-key chunk of Vt: 64 head-dim rows x 64 B
constexpr int P_TILE = BQ * 64;                  // one (buffer, part) P chunk: 128 rows x 32 keys
constexpr int Q_BYTES = 4 * Q_TILE, K_BYTES = 4 * K_TILE, V_BYTES = 2 * (KMAX / KC) * V_TILE, P_BYTES = 4 * P_TILE;
constexpr int BAR_BYTES = 256, TOFF_BYTES = (MAX_GROUPS + 1) * 4;
constexpr int SMEM_BYTES = Q_BYTES + K_BYTES + V_BYTES + P_BYTES + BAR_BYTES + TOFF_BYTES + 1024;
static_assert(SMEM_BYTES <= 232448, "exceeds 227 KB of shared memory");
constexpr int TMEM_COLS = 512, S_COL = 0, O_COL = 256;

struct Params {
    const int* seq_off; int q_mul;             // queries of wireframe i: rows [seq_off[i]*q_mul, seq_off[i+1]*q_mul)
    const int* row_off; const int* vlen;       // keys of wireframe i: cache rows [row_off[i], +vlen[i])
    const int* colp_off;                       // first column of wireframe i in the transposed value cache (multiple of 32)
    int n_groups, n_heads;
    int layer_col;                             // first column of this layer in the cache rows (layer * E)
    int total_items;                           // n_heads * sum_i ceil(queries_i / 128)
    uint16_t* Os; long long os_stride; int ldo;   // fp16x2 output [2][rows][ldo]
    const int* stop;
};

__device__ __forceinline__ uint32_t idesc_f16(int n) {     // kind::f16, fp16 A/B, fp32 D, K-major A and B, M = 128
    return (1u << 4) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(BQ >> 4) << 24);
}

struct Cursor {                                 // position in the item order (wireframe, head, tile), tile fastest
    int wf, head, tile, tiles;
    __device__ __forceinline__ void seek(const int* toff, int n_groups, int n_heads, int idx) {
        wf = 0;
        while (wf + 1 < n_groups && idx >= n_heads * toff[wf + 1]) ++wf;
        tiles = toff[wf + 1] - toff[wf];
        const int rem = idx - n_heads * toff[wf];
        head = rem / tiles; tile = rem - head * tiles;
    }
    __device__ __forceinline__ void next(const int* toff, int n_groups, int n_heads) {
        if (++tile < tiles) return;
        tile = 0;
        if (++head < n_heads) return;
        head = 0;
        do { ++wf; tiles = (wf < n_groups) ? toff[wf + 1] - toff[wf] : 1; } while (wf < n_groups && tiles == 0);
    }
};

__global__ void __launch_bounds__(NUM_THREADS, 1)
attn_x_kernel(const __grid_constant__ CUtensorMap mapQ, const __grid_constant__ CUtensorMap mapK,
              const __grid_constant__ CUtensorMap mapVt, const Params p) {
    using namespace tc;
    if (p.stop != nullptr && *p.stop != 0) return;

    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
    const uint32_t q_s = smem_base, k_s = q_s + Q_BYTES, v_s = k_s + K_BYTES, p_s = v_s + V_BYTES;
    const uint32_t bar_base = p_s + P_BYTES;
    uint8_t* p_gen = smem_gen + (p_s - smem_base);
    int* toff = reinterpret_cast<int*>(smem_gen + (bar_base - smem_base) + BAR_BYTES);
    // barriers
    const uint32_t q_full = bar_base, q_empty = bar_base + 8, kv_full = bar_base + 16, kv_empty = bar_base + 24;
    const uint32_t s_full = bar_base + 32, o_full = bar_base + 40, tmem_empty = bar_base + 48;
    auto p_full = [&](uint32_t b) { return bar_base + 56 + 8 * b; };
    auto p_empty = [&](uint32_t b) { return bar_base + 72 + 8 * b; };
    const uint32_t tmem_slot = bar_base + 96;
    uint32_t* tmem_slot_ptr = reinterpret_cast<uint32_t*>(smem_gen + (tmem_slot - smem_base));

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int G = p.n_groups, H = p.n_heads;

    // tiles per wireframe -> exclusive prefix in shared memory
    for (int i = threadIdx.x; i < G; i += NUM_THREADS) {
        const long long nq = ((long long)p.seq_off[i + 1] - p.seq_off[i]) * p.q_mul;
        toff[i + 1] = (int)((nq + BQ - 1) / BQ);
    }
    if (threadIdx.x == 0) {
        mbar_init(q_full, 1); mbar_init(q_empty, 1); mbar_init(kv_full, 1); mbar_init(kv_empty, 1);
        mbar_init(s_full, 1); mbar_init(o_full, 1); mbar_init(tmem_empty, 128);
        mbar_init(p_full(0), 128); mbar_init(p_full(1), 128); mbar_init(p_empty(0), 1); mbar_init(p_empty(1), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) tmem_alloc(tmem_slot, TMEM_COLS);
    __syncthreads();
    if (threadIdx.x == 0) {
        int acc = 0; toff[0] = 0;
        for (int i = 1; i <= G; ++i) { acc += toff[i]; toff[i] = acc; }
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot_ptr;

    const int item0 = (int)((long long)blockIdx.x * p.total_items / gridDim.x);
    const int item1 = (int)((long long)(blockIdx.x + 1) * p.total_items / gridDim.x);

    if (warp == 0) {
        if (lane == 0 && item0 < item1) {
            // ===== TMA producer =====
            Cursor c; c.seek(toff, G, H, item0);
            int cur_wf = -1, cur_head = -1; uint32_t n_kv = 0, n_q = 0;
            for (int it = item0; it < item1; ++it, c.next(toff, G, H)) {
                if (c.wf != cur_wf || c.head != cur_head) {
                    cur_wf = c.wf; cur_head = c.head;
                    mbar_wait(kv_empty, (n_kv & 1u) ^ 1u);                 // every MMA that read the previous K / Vt has retired
                    const int nb = (p.vlen[c.wf] + KC - 1) / KC;
                    mbar_expect_tx(kv_full, (uint32_t)nb * (4u * 2048u + 2u * V_TILE));
                    const int col = p.layer_col + c.head * 64, krow = p.row_off[c.wf], vcol = p.colp_off[c.wf];
                    for (int part = 0; part < 2; ++part)
                        for (int kch = 0; kch < 2; ++kch)
                            for (int b = 0; b < nb; ++b)
                                tma_load_3d(k_s + (part * 2 + kch) * K_TILE + b * 2048, &mapK, kv_full, col + kch * 32, krow + b * KC, part);
                    for (int part = 0; part < 2; ++part)
                        for (int b = 0; b < nb; ++b)
                            tma_load_3d(v_s + (part * (KMAX / KC) + b) * V_TILE, &mapVt, kv_full, vcol + b * KC, col, part);
                    ++n_kv;
                }
                mbar_wait(q_empty, (n_q & 1u) ^ 1u);
                mbar_expect_tx(q_full, Q_BYTES);
                const int qrow = p.seq_off[c.wf] * p.q_mul + c.tile * BQ;
                for (int part = 0; part < 2; ++part)
                    for (int kch = 0; kch < 2; ++kch)
                        tma_load_3d(q_s + (part * 2 + kch) * Q_TILE, &mapQ, q_full, c.head * 64 + kch * 32, qrow, part);
                ++n_q;
            }
        }
    } else if (warp == 1) {
        if (lane == 0 && item0 < item1) {
            // ===== MMA issuer =====
            Cursor c; c.seek(toff, G, H, item0);
            int cur_wf = -1, cur_head = -1; uint32_t n_kv = 0, n_it = 0, n_pc = 0;
            const uint32_t d_s = tmem_base + S_COL, d_o = tmem_base + O_COL;
            const uint32_t idesc_o = idesc_f16(64);
            constexpr int pa[3] = {1, 0, 0}, pb[3] = {0, 1, 0};            // (lo,hi), (hi,lo), (hi,hi): small products first
            for (int it = item0; it < item1; ++it) {
                const int wf = c.wf, head = c.head;
                if (wf != cur_wf || head != cur_head) {
                    cur_wf = wf; cur_head = head;
                    mbar_wait(kv_full, n_kv & 1u);
                    ++n_kv;
                }
                const int lv = p.vlen[wf];
                const int n16 = (lv + 15) & ~15, nb = (lv + KC - 1) / KC;
                mbar_wait(q_full, n_it & 1u);
                mbar_wait(tmem_empty, (n_it & 1u) ^ 1u);                   // S and O of the previous item have been read out
                tc_fence_after();
                const uint32_t idesc_s = idesc_f16(n16);
                uint32_t acc = 0;
#pragma unroll
                for (int kch = 0; kch < 2; ++kch)
#pragma unroll
                    for (int ks = 0; ks < 2; ++ks)
#pragma unroll
                        for (int q = 0; q < 3; ++q) {
                            const uint64_t da = make_smem_desc(q_s + (pa[q] * 2 + kch) * Q_TILE) + (uint64_t)(2 * ks);
                            const uint64_t db = make_smem_desc(k_s + (pb[q] * 2 + kch) * K_TILE) + (uint64_t)(2 * ks);
                            umma_bf16(d_s, da, db, idesc_s, acc);
                            acc = 1;
                        }
                umma_commit(q_empty);
                umma_commit(s_full);
                c.next(toff, G, H);
                const bool last_of_group = (it + 1 == item1) || c.wf != wf || c.head != head;
                acc = 0;
                for (int b = 0; b < nb; ++b, ++n_pc) {
                    const uint32_t buf = n_pc & 1u;
                    mbar_wait(p_full(buf), (n_pc >> 1) & 1u);
                    tc_fence_after();
#pragma unroll
                    for (int ks = 0; ks < 2; ++ks)
#pragma unroll
                        for (int q = 0; q < 3; ++q) {
                            const uint64_t da = make_smem_desc(p_s + (buf * 2 + pa[q]) * P_TILE) + (uint64_t)(2 * ks);
                            const uint64_t db = make_smem_desc(v_s + (pb[q] * (KMAX / KC) + b) * V_TILE) + (uint64_t)(2 * ks);
                            umma_bf16(d_o, da, db, idesc_o, acc);
                            acc = 1;
                        }
                    umma_commit(p_empty(buf));
                }
                umma_commit(o_full);
                if (last_of_group) umma_commit(kv_empty);
                ++n_it;
            }
        }
    } else {
        // ===== softmax + epilogue warps: thread = query row =====
        const int q = warp & 3;                               // TMEM lane quarter this warp may access
        const int row = q * 32 + lane;
        const uint32_t t_lane = tmem_base + ((uint32_t)(q * 32) << 16);
        constexpr float kScale = 0.125f * 1.4426950408889634f;
        // per-warp output staging: four 2 KB pieces that coincide with this warp's own rows of the four P tiles
        auto piece = [&](int k) { return p_gen + k * P_TILE + q * 2048; };
        Cursor c; c.seek(toff, G, H, item0 < item1 ? item0 : 0);
        uint32_t n_it = 0, n_pc = 0;
        for (int it = item0; it < item1; ++it, c.next(toff, G, H), ++n_it) {
            const int lv = p.vlen[c.wf], nb = (lv + KC - 1) / KC;
            const long long q_end = (long long)p.seq_off[c.wf + 1] * p.q_mul;
            const long long qrow0 = (long long)p.seq_off[c.wf] * p.q_mul + (long long)c.tile * BQ;
            mbar_wait(s_full, n_it & 1u);
            tc_fence_after();
            // ---- pass 1: row maximum over the valid keys ----
            float mx = -INFINITY;
            for (int b = 0; b < nb; ++b) {
                uint32_t v[32];
                tmem_ld32(t_lane + S_COL + b * KC, v);
                tmem_ld_wait();
                const int rem = lv - b * KC;
#pragma unroll
                for (int i = 0; i < 32; ++i) if (i < rem) mx = fmaxf(mx, __uint_as_float(v[i]));
            }
            const float bias = 12.0f - mx * kScale;           // p' = 2^(s*kScale - m' + 12) = 4096 * exp((s - m)/8)
            float lsum = 0.f;
            // ---- pass 2: probabilities per 32-key chunk -> fp16x2 -> shared memory (A operand of O += P V) ----
            for (int b = 0; b < nb; ++b, ++n_pc) {
                uint32_t v[32];
                tmem_ld32(t_lane + S_COL + b * KC, v);
                tmem_ld_wait();
                const uint32_t buf = n_pc & 1u;
                const int rem = lv - b * KC;
                uint32_t hw[16], lw[16];
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    float p0 = ex2_approx(fmaf(__uint_as_float(v[2 * i]), kScale, bias));
                    float p1 = ex2_approx(fmaf(__uint_as_float(v[2 * i + 1]), kScale, bias));
                    p0 = (2 * i < rem) ? p0 : 0.f;
                    p1 = (2 * i + 1 < rem) ? p1 : 0.f;
                    lsum += p0 + p1;
                    split_pair(p0, p1, hw[i], lw[i]);
                }
                mbar_wait(p_empty(buf), ((n_pc >> 1) & 1u) ^ 1u);         // the MMAs that read this buffer two chunks ago have retired
                uint8_t* ph = p_gen + (buf * 2 + 0) * P_TILE + row * 64;
                uint8_t* pl = p_gen + (buf * 2 + 1) * P_TILE + row * 64;
#pragma unroll
                for (int cc = 0; cc < 4; ++cc) {                          // 16-byte chunk cc lives at cc ^ ((row >> 1) & 3) (SWIZZLE_64B)
                    const int sw = (cc ^ ((row >> 1) & 3)) << 4;
                    *reinterpret_cast<uint4*>(ph + sw) = make_uint4(hw[4 * cc], hw[4 * cc + 1], hw[4 * cc + 2], hw[4 * cc + 3]);
                    *reinterpret_cast<uint4*>(pl + sw) = make_uint4(lw[4 * cc], lw[4 * cc + 1], lw[4 * cc + 2], lw[4 * cc + 3]);
                }
                fence_proxy_async_smem();
                mbar_arrive(p_full(buf));
            }
            // ---- epilogue: O / l -> fp16x2 -> global ----
            mbar_wait(o_full, n_it & 1u);
            tc_fence_after();
            uint32_t o0[32], o1[32];
            tmem_ld32(t_lane + O_COL, o0);
            tmem_ld32(t_lane + O_COL + 32, o1);
            tmem_ld_wait();
            tc_fence_before();
            mbar_arrive(tmem_empty);
            const float inv = 1.0f / lsum;
            uint8_t* sh = piece(lane >> 4) + (lane & 15) * 128;           // hi row of this thread
            uint8_t* sl = piece(2 + (lane >> 4)) + (lane & 15) * 128;     // lo row
#pragma unroll
            for (int cc = 0; cc < 8; ++cc) {
                uint32_t hw[4], lw[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int i = 8 * cc + 2 * u;
                    const float x0 = __uint_as_float(i < 32 ? o0[i & 31] : o1[i & 31]) * inv;
                    const float x1 = __uint_as_float(i < 32 ? o0[(i + 1) & 31] : o1[(i + 1) & 31]) * inv;
                    split_pair(x0, x1, hw[u], lw[u]);
                }
                const int sw = (cc ^ (lane & 7)) << 4;
                *reinterpret_cast<uint4*>(sh + sw) = make_uint4(hw[0], hw[1], hw[2], hw[3]);
                *reinterpret_cast<uint4*>(sl + sw) = make_uint4(lw[0], lw[1], lw[2], lw[3]);
            }
            __syncwarp();
#pragma unroll
            for (int part = 0; part < 2; ++part)
#pragma unroll
                for (int r4 = 0; r4 < 8; ++r4) {
                    const int r = r4 * 4 + (lane >> 3), ch = lane & 7;    // 8 lanes store one 128-byte row
                    const uint4 val = *reinterpret_cast<const uint4*>(piece(2 * part + (r >> 4)) + (r & 15) * 128 + ((ch ^ (r & 7)) << 4));
                    const long long grow = qrow0 + q * 32 + r;
                    if (grow < q_end)
                        *reinterpret_cast<uint4*>(p.Os + (size_t)part * p.os_stride + (size_t)grow * p.ldo + c.head * 64 + ch * 8) = val;
                }
            __syncwarp();
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, TMEM_COLS);
}

// Transposed fp16x2 value cache: dst[part][col][colp_off[g] + j] = split(src[row_off[g] + j][col]); padding columns stay zero
// (the buffer is zero-filled first).  grid (ceil(ld / 32), n_groups), block (32, 8).
__global__ void build_vt_kernel(const float* __restrict__ src, int ld, const int* __restrict__ row_off, const int* __restrict__ vlen,
                                const int* __restrict__ colp_off, uint16_t* __restrict__ dst, long long rp, int* ovf) {
    __shared__ float tile[32][33];
    const int g = blockIdx.y, c0 = blockIdx.x * 32;
    const int r0 = row_off[g], n = vlen[g], cp = colp_off[g];
    const long long part_stride = (long long)ld * rp;
    for (int j0 = 0; j0 < n; j0 += 32) {
        for (int jj = threadIdx.y; jj < 32; jj += 8) {
            const int j = j0 + jj, col = c0 + threadIdx.x;
            tile[jj][threadIdx.x] = (j < n && col < ld) ? src[(size_t)(r0 + j) * ld + col] : 0.f;
        }
        __syncthreads();
        for (int cc = threadIdx.y; cc < 32; cc += 8) {
            const int j = j0 + threadIdx.x, col = c0 + cc;
            if (j < n && col < ld) store_split1(dst + (size_t)col * rp + cp + j, part_stride, tile[threadIdx.x][cc], 2, ovf);
        }
        __syncthreads();
    }
}

}  // namespace ax
}  // namespace ffb
