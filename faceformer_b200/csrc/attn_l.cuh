// attn_l.cuh -- encoder self-attention over MANY keys on tcgen05 + TMEM + TMA (sm_100a only).
//
//   O = softmax(Q K^T / 8) V   per (wireframe, head); queries = keys = the wireframe's memory rows (transformer.py:164-176 with the
//   key-padding mask realised as "only valid rows exist").  attn_x.cuh keeps all keys of a group in shared memory (<= 256); the
//   2048-edge wireframes of BASELINE.json configs[4] have 2052, and half of that encoder's FLOPs are attention.  Here K / V STREAM
//   through shared memory in 64-key blocks and the softmax is exact in two passes over the keys:
//     pass 1   row maximum from S~ = Q_hi K_hi^T (ONE fp16 product: softmax is shift-invariant, so any m within a fraction of the true
//              maximum serves; only the hi halves of K are loaded)                          4 MMAs M128 N64 K16 per block and tile
//     pass 2   S = Q K^T (3 products) -> p = 2^((s - m) * log2(e) / 8 + 12) -> fp16x2 -> smem;  O += P V         12 + 12 MMAs
//   No running-maximum rescale of O is needed.  O is drained from TMEM into fp32 registers every DRAIN blocks so that accumulation
//   chains stay short (the tensor core accumulates with truncation, gemm_tc.cuh).
//
// A work item is (wireframe, head, PAIR of 128-query tiles): both tiles share every K / V block that streams through shared memory,
// which halves the L2 -> SM traffic per query (the first version, one tile per item, was L2-bound: 48 KB per 128 x 64 block of
// scores) and gives the tensor pipe two independent softmax warpgroups to alternate between.
//
// Operands are the fp16x2 splits the QKV GEMM epilogue wrote ([2][rows][3E]); every product is  lo*hi + hi*lo + hi*hi.
// Layouts are those of attn_x.cuh: Q / K as K-major 64-byte rows (SWIZZLE_64B, 32 head dims per tile), V rows as an MN-major
// B operand (SWIZZLE_128B), P written by the softmax threads as the SWIZZLE_64B image of a K-major A tile.
//
// One persistent CTA per SM walks a contiguous range of work items.  Roles (384 threads):
//   warp 0      TMA producer: both Q tiles once per item; K blocks twice (pass 1: hi halves only) through a 3-slot ring; V blocks
//               through a 2-slot ring
//   warp 1      TMEM allocator + the S issuer (one thread): per block S for both tiles, up to two blocks ahead of the softmax
//   warp 2      the P V issuer (one thread): O += P V as the softmax publishes P; its waits never hold back the S issuer
//               (with ONE issuer thread the softmax warps spent 40 % of their time waiting for S: profiles/ncu_attn_long_r2_summary.md)
//   warps 4-7   softmax + epilogue of tile 0, warps 8-11 of tile 1: one query row per thread (TMEM lane = row)
// TMEM: per tile two S buffers of 64 columns and one O of 64 columns.
#pragma once
#include "attn_x.cuh"

namespace ffb {
namespace al {

#ifndef FFB_AL_DRAIN
#define FFB_AL_DRAIN 8
#endif
constexpr int BQ = 128, KB = 64, KC = 32, NUM_THREADS = 384, NKS = 3, NVS = 2, DRAIN = FFB_AL_DRAIN;
constexpr int Q_TILE = BQ * 64;                    // one (part, head-dim chunk) Q tile: 128 rows x 64 B
constexpr int K_TILE = KB * 64;                    // one (part, head-dim chunk) K tile of a block: 64 rows x 64 B
constexpr int V_TILE = KC * 128;                   // 32 keys x 128 B
constexpr int P_TILE = BQ * 64;                    // one (part, 32-key chunk) P tile: 128 rows x 64 B
constexpr int Q_BYTES = 4 * Q_TILE, K_SLOT = 4 * K_TILE, V_SLOT = 4 * V_TILE, P_BUF = 4 * P_TILE;
constexpr int SMEM_BYTES = 2 * Q_BYTES + NKS * K_SLOT + NVS * V_SLOT + 2 * P_BUF + 256 + 1024;
static_assert(SMEM_BYTES <= 232448, "exceeds 227 KB of shared memory");
constexpr int TMEM_COLS = 512;                     // tile w: S0 [w*256, +64), S1 [w*256 + 64, +64), O [w*256 + 128, +64)

struct Params {
    const int* tile_off;                           // [n_groups + 1] exclusive prefix of ceil(ceil(vlen / 128) / 2): tile PAIRS per group
    const int* row_off; const int* vlen;           // rows of group g: [row_off[g], row_off[g] + vlen[g])
    int n_groups, n_heads;
    int q_col, k_col, v_col;                       // first column of head 0 in the Q / K / V rows
    int total_items;
    uint16_t* Os; long long os_stride; int ldo;    // fp16x2 output [2][rows][ldo]
};

struct Item { int q_row0, q_rows, k_row0, nk, head; };   // q_rows: valid query rows of the pair (<= 256)

__device__ __forceinline__ void get_item(const Params& p, int idx, Item& it) {
    const int H = p.n_heads, t = idx / H;
    int lo = 0, hi = p.n_groups - 1;               // largest g with tile_off[g] <= t
    while (lo < hi) { const int mid = (lo + hi + 1) >> 1; if (p.tile_off[mid] <= t) lo = mid; else hi = mid - 1; }
    const int g = lo, pairs = p.tile_off[g + 1] - p.tile_off[g];
    const int rem = idx - H * p.tile_off[g];
    it.head = rem / pairs;
    const int pr = rem - it.head * pairs;
    it.k_row0 = p.row_off[g]; it.nk = p.vlen[g];
    it.q_row0 = it.k_row0 + pr * 2 * BQ; it.q_rows = min(2 * BQ, it.nk - pr * 2 * BQ);
}

__global__ void __launch_bounds__(NUM_THREADS, 1)
attn_long_kernel(const __grid_constant__ CUtensorMap mapQ, const __grid_constant__ CUtensorMap mapK,
                 const __grid_constant__ CUtensorMap mapV, const Params p) {
    using namespace tc;
    using ax::idesc_f16;
    using ax::make_smem_desc_mn128;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
    const uint32_t q_s = smem_base, k_s = q_s + 2 * Q_BYTES, v_s = k_s + NKS * K_SLOT, p_s = v_s + NVS * V_SLOT;
    const uint32_t bar_base = p_s + 2 * P_BUF;
    uint8_t* p_gen = smem_gen + (p_s - smem_base);
    const uint32_t q_full = bar_base, q_empty = bar_base + 8;
    auto k_full = [&](uint32_t s) { return bar_base + 16 + 8 * s; };              // 3
    auto k_empty = [&](uint32_t s) { return bar_base + 40 + 8 * s; };             // 3
    auto v_full = [&](uint32_t s) { return bar_base + 64 + 8 * s; };              // 2
    auto v_empty = [&](uint32_t s) { return bar_base + 80 + 8 * s; };             // 2
    auto s_full = [&](uint32_t w, uint32_t b) { return bar_base + 96 + 8 * (2 * w + b); };     // 4
    auto s_empty = [&](uint32_t w, uint32_t b) { return bar_base + 128 + 8 * (2 * w + b); };   // 4
    auto p_full = [&](uint32_t w) { return bar_base + 160 + 8 * w; };
    auto p_empty = [&](uint32_t w) { return bar_base + 176 + 8 * w; };
    auto o_full = [&](uint32_t w) { return bar_base + 192 + 8 * w; };
    auto o_empty = [&](uint32_t w) { return bar_base + 208 + 8 * w; };
    const uint32_t tmem_slot = bar_base + 224;
    uint32_t* tmem_slot_ptr = reinterpret_cast<uint32_t*>(smem_gen + (tmem_slot - smem_base));

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        mbar_init(q_full, 1); mbar_init(q_empty, 1);
        for (uint32_t s = 0; s < NKS; ++s) { mbar_init(k_full(s), 1); mbar_init(k_empty(s), 1); }
        for (uint32_t s = 0; s < NVS; ++s) { mbar_init(v_full(s), 1); mbar_init(v_empty(s), 1); }
        for (uint32_t w = 0; w < 2; ++w) {
            for (uint32_t b = 0; b < 2; ++b) { mbar_init(s_full(w, b), 1); mbar_init(s_empty(w, b), 128); }
            mbar_init(p_full(w), 128); mbar_init(p_empty(w), 1); mbar_init(o_full(w), 1); mbar_init(o_empty(w), 128);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) tmem_alloc(tmem_slot, TMEM_COLS);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot_ptr;

    const int item0 = (int)((long long)blockIdx.x * p.total_items / gridDim.x);
    const int n_items = (int)((long long)(blockIdx.x + 1) * p.total_items / gridDim.x) - item0;

    if (warp < 4) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 72;");                      // producer / issuer warpgroup gives its registers ...
    if (warp == 0) {
        if (lane == 0) {
            // ===== TMA producer =====
            uint32_t nk_ld = 0, nv_ld = 0;                   // K / V blocks loaded so far (ring positions)
            for (int j = 0; j < n_items; ++j) {
                Item it; get_item(p, item0 + j, it);
                const int nb = (it.nk + KB - 1) / KB;
                mbar_wait(q_empty, ((uint32_t)j & 1u) ^ 1u);
                mbar_expect_tx(q_full, 2 * Q_BYTES);
                for (int w = 0; w < 2; ++w)
                    for (int part = 0; part < 2; ++part)
                        for (int kch = 0; kch < 2; ++kch)
                            tma_load_3d(q_s + w * Q_BYTES + (part * 2 + kch) * Q_TILE, &mapQ, q_full, p.q_col + it.head * 64 + kch * 32,
                                        it.q_row0 + w * BQ, part);
                for (int pass = 0; pass < 2; ++pass)
                    for (int b = 0; b < nb; ++b) {
                        const uint32_t ks = nk_ld % NKS;
                        const int parts = pass == 0 ? 1 : 2;                      // pass 1 multiplies the hi halves only
                        mbar_wait(k_empty(ks), ((nk_ld / NKS) & 1u) ^ 1u);
                        mbar_expect_tx(k_full(ks), (uint32_t)parts * 2u * K_TILE);
                        for (int part = 0; part < parts; ++part)
                            for (int kch = 0; kch < 2; ++kch)
                                for (int h2 = 0; h2 < 2; ++h2)
                                    tma_load_3d(k_s + ks * K_SLOT + (part * 2 + kch) * K_TILE + h2 * 2048, &mapK, k_full(ks),
                                                p.k_col + it.head * 64 + kch * 32, it.k_row0 + b * KB + h2 * KC, part);
                        ++nk_ld;
                        if (pass == 1) {
                            const uint32_t vs = nv_ld % NVS;
                            mbar_wait(v_empty(vs), ((nv_ld / NVS) & 1u) ^ 1u);
                            mbar_expect_tx(v_full(vs), V_SLOT);
                            for (int part = 0; part < 2; ++part)
                                for (int c = 0; c < 2; ++c)
                                    tma_load_3d(v_s + vs * V_SLOT + (part * 2 + c) * V_TILE, &mapV, v_full(vs), p.v_col + it.head * 64,
                                                it.k_row0 + b * KB + c * KC, part);
                            ++nv_ld;
                        }
                    }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            // ===== S issuer: runs ahead of the softmax by up to two blocks (two S buffers per tile) =====
            constexpr int pa[3] = {1, 0, 0}, pb[3] = {0, 1, 0};            // (lo,hi), (hi,lo), (hi,hi): small products first
            const uint32_t idesc_s = idesc_f16(KB, 0);
            uint32_t nk_use = 0, ns = 0;                                   // K blocks and S buffers (per tile) consumed
            // S of BOTH tiles against one K block; hi_only: the single product of pass 1
            auto issue_s = [&](bool hi_only) {
                const uint32_t ks = nk_use % NKS, sb = ns & 1u;
                mbar_wait(k_full(ks), (nk_use / NKS) & 1u);
                const uint32_t kb = k_s + ks * K_SLOT;
#pragma unroll
                for (uint32_t w = 0; w < 2; ++w) {
                    mbar_wait(s_empty(w, sb), ((ns >> 1) & 1u) ^ 1u);
                    tc_fence_after();
                    const uint32_t d_s = tmem_base + w * 256u + sb * 64u, qb = q_s + w * Q_BYTES;
                    uint32_t acc = 0;
#pragma unroll
                    for (int kch = 0; kch < 2; ++kch)
#pragma unroll
                        for (int s2 = 0; s2 < 2; ++s2)
#pragma unroll
                            for (int q = 0; q < 3; ++q) {
                                if (hi_only && q != 2) continue;
                                const uint64_t da = make_smem_desc(qb + (pa[q] * 2 + kch) * Q_TILE) + (uint64_t)(2 * s2);
                                const uint64_t db = make_smem_desc(kb + (pb[q] * 2 + kch) * K_TILE) + (uint64_t)(2 * s2);
                                umma_bf16(d_s, da, db, idesc_s, acc);
                                acc = 1;
                            }
                    if (w == 1) umma_commit(k_empty(ks));
                    umma_commit(s_full(w, sb));
                }
                ++nk_use; ++ns;
            };
            for (int j = 0; j < n_items; ++j) {
                Item it; get_item(p, item0 + j, it);
                const int nb = (it.nk + KB - 1) / KB;
                mbar_wait(q_full, (uint32_t)j & 1u);
                for (int b = 0; b < nb; ++b) issue_s(true);                // pass 1
                for (int b = 0; b < nb; ++b) issue_s(false);               // pass 2
                umma_commit(q_empty);                                      // every S product of this item has been issued
            }
        }
    } else if (warp == 2) {
        if (lane == 0) {
            // ===== P V issuer: O += P(b) V(b) as the softmax warpgroups publish P (its waits never hold back the S issuer) =====
            constexpr int pa[3] = {1, 0, 0}, pb[3] = {0, 1, 0};
            const uint32_t idesc_o = idesc_f16(64, 1);
            uint32_t nv_use = 0, npv = 0, n_og = 0;                        // V blocks, P uses (per tile), O drain groups consumed
            for (int j = 0; j < n_items; ++j) {
                Item it; get_item(p, item0 + j, it);
                const int nb = (it.nk + KB - 1) / KB;
                for (int bb = 0; bb < nb; ++bb) {
                    const uint32_t vs = nv_use % NVS;
                    const bool first = (bb % DRAIN) == 0;                  // first block of a drain group: O restarts from zero
                    const bool last = (bb % DRAIN) == DRAIN - 1 || bb == nb - 1;
                    mbar_wait(v_full(vs), (nv_use / NVS) & 1u);
#pragma unroll
                    for (uint32_t w = 0; w < 2; ++w) {
                        mbar_wait(p_full(w), npv & 1u);
                        if (first) mbar_wait(o_empty(w), (n_og & 1u) ^ 1u);       // the softmax threads have read the previous group's O
                        tc_fence_after();
                        const uint32_t d_o = tmem_base + w * 256u + 128u;
                        uint32_t acc = first ? 0u : 1u;
#pragma unroll
                        for (int c = 0; c < 2; ++c)
#pragma unroll
                            for (int s2 = 0; s2 < 2; ++s2)
#pragma unroll
                                for (int q = 0; q < 3; ++q) {
                                    const uint64_t da = make_smem_desc(p_s + w * P_BUF + (pa[q] * 2 + c) * P_TILE) + (uint64_t)(2 * s2);
                                    const uint64_t db = make_smem_desc_mn128(v_s + vs * V_SLOT + (pb[q] * 2 + c) * V_TILE + s2 * 2048u);
                                    umma_bf16(d_o, da, db, idesc_o, acc);
                                    acc = 1;
                                }
                        umma_commit(p_empty(w));
                        if (w == 1) umma_commit(v_empty(vs));
                        if (last) umma_commit(o_full(w));
                    }
                    ++npv; ++nv_use;
                    if (last) ++n_og;
                }
            }
        }
    }
    } else {
        asm volatile("setmaxnreg.inc.sync.aligned.u32 216;");                 // ... to the two softmax warpgroups (64 fp32 O accumulators per row)
        // ===== softmax + epilogue: warpgroup w owns tile w of the pair; thread = query row =====
        const uint32_t w = (uint32_t)(warp - 4) >> 2;
        const int q = warp & 3;                               // TMEM lane quarter this warp may access
        const int row = q * 32 + lane;
        const uint32_t t_row = tmem_base + ((uint32_t)(q * 32) << 16) + w * 256u;
        constexpr float kScale = 0.125f * 1.4426950408889634f;
        const int rsw = (row >> 1) & 3;                       // 16-byte chunk cc of a P row lives at cc ^ ((row >> 1) & 3) (SWIZZLE_64B)
        uint8_t* pw = p_gen + w * P_BUF;
        uint32_t ns = 0, npv = 0, n_og = 0;
        for (int j = 0; j < n_items; ++j) {
            Item it; get_item(p, item0 + j, it);
            const int nb = (it.nk + KB - 1) / KB;
            // ---- pass 1: row maximum (of the hi x hi product: within a fraction of the true maximum, which is all the shift needs) ----
            float mx = -INFINITY;
            for (int b = 0; b < nb; ++b, ++ns) {
                const uint32_t sb = ns & 1u;
                mbar_wait(s_full(w, sb), (ns >> 1) & 1u);
                tc_fence_after();
                uint32_t v0[32], v1[32];
                tmem_ld32(t_row + sb * 64u, v0);
                tmem_ld32(t_row + sb * 64u + 32u, v1);
                tmem_ld_wait();
                tc_fence_before();
                mbar_arrive(s_empty(w, sb));
                const int nv = min(KB, it.nk - b * KB);       // valid keys of this block (the rows behind them belong to the next wireframe)
                if (nv == KB) {
                    float m0 = mx, m1 = -INFINITY;
#pragma unroll
                    for (int i = 0; i < 32; i += 2) {
                        m0 = fmaxf(m0, fmaxf(__uint_as_float(v0[i]), __uint_as_float(v0[i + 1])));
                        m1 = fmaxf(m1, fmaxf(__uint_as_float(v1[i]), __uint_as_float(v1[i + 1])));
                    }
                    mx = fmaxf(m0, m1);
                } else {
#pragma unroll
                    for (int i = 0; i < 32; ++i) {
                        mx = fmaxf(mx, i < nv ? __uint_as_float(v0[i]) : -INFINITY);
                        mx = fmaxf(mx, 32 + i < nv ? __uint_as_float(v1[i]) : -INFINITY);
                    }
                }
            }
            // p' = 2^((s - m~) * kScale + 11.5): m~ is the maximum of the single-product scores, at most a few 2^-11 relative below the
            // true maximum, so p' <= 2^11.5 * e^(small) stays far inside fp16's range while its low half stays out of the subnormals
            const float bias = 11.5f - mx * kScale;
            float lsum = 0.f;
            float oacc[64];
#pragma unroll
            for (int i = 0; i < 64; ++i) oacc[i] = 0.f;
            auto drain_o = [&]() {                            // O of one drain group -> register accumulators (round-to-nearest adds)
                mbar_wait(o_full(w), n_og & 1u);
                tc_fence_after();
                uint32_t o0[32], o1[32];
                tmem_ld32(t_row + 128u, o0);
                tmem_ld32(t_row + 160u, o1);
                tmem_ld_wait();
                tc_fence_before();
                mbar_arrive(o_empty(w));
                ++n_og;
#pragma unroll
                for (int i = 0; i < 32; ++i) { oacc[i] += __uint_as_float(o0[i]); oacc[32 + i] += __uint_as_float(o1[i]); }
            };
            // ---- pass 2: probabilities per 64-key block -> fp16x2 -> shared memory (A operand of O += P V) ----
            for (int b = 0; b < nb; ++b, ++ns, ++npv) {
                const uint32_t sb = ns & 1u;
                mbar_wait(s_full(w, sb), (ns >> 1) & 1u);
                tc_fence_after();
                uint32_t v0[32], v1[32];
                tmem_ld32(t_row + sb * 64u, v0);
                tmem_ld32(t_row + sb * 64u + 32u, v1);
                tmem_ld_wait();
                tc_fence_before();
                mbar_arrive(s_empty(w, sb));
                const int nv = min(KB, it.nk - b * KB);
                uint32_t hw[32], lw[32];
                float ls0 = 0.f, ls1 = 0.f;
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    float p0 = ex2_approx(fmaf(__uint_as_float(v0[2 * i]), kScale, bias));
                    float p1 = ex2_approx(fmaf(__uint_as_float(v0[2 * i + 1]), kScale, bias));
                    float p2 = ex2_approx(fmaf(__uint_as_float(v1[2 * i]), kScale, bias));
                    float p3 = ex2_approx(fmaf(__uint_as_float(v1[2 * i + 1]), kScale, bias));
                    if (nv < KB) {
                        p0 = (2 * i < nv) ? p0 : 0.f; p1 = (2 * i + 1 < nv) ? p1 : 0.f;
                        p2 = (32 + 2 * i < nv) ? p2 : 0.f; p3 = (33 + 2 * i < nv) ? p3 : 0.f;
                    }
                    ls0 += p0 + p1; ls1 += p2 + p3;
                    split_pair(p0, p1, hw[i], lw[i]);
                    split_pair(p2, p3, hw[16 + i], lw[16 + i]);
                }
                lsum += ls0 + ls1;
                // the drain of the previous group sits between this block's exponentials and its publication: the MMA thread is then at most
                // one block behind and the wait is short
                if (b > 0 && (b % DRAIN) == 0) drain_o();
                mbar_wait(p_empty(w), (npv & 1u) ^ 1u);       // O += P V of the previous block has consumed the (single) P buffer
#pragma unroll
                for (int c = 0; c < 2; ++c) {                 // 32-key chunk c: tiles (hi, c) and (lo, c)
                    uint8_t* ph = pw + (0 * 2 + c) * P_TILE + row * 64;
                    uint8_t* pl = pw + (1 * 2 + c) * P_TILE + row * 64;
#pragma unroll
                    for (int cc = 0; cc < 4; ++cc) {
                        const int sw = (cc ^ rsw) << 4;
                        *reinterpret_cast<uint4*>(ph + sw) = make_uint4(hw[16 * c + 4 * cc], hw[16 * c + 4 * cc + 1], hw[16 * c + 4 * cc + 2], hw[16 * c + 4 * cc + 3]);
                        *reinterpret_cast<uint4*>(pl + sw) = make_uint4(lw[16 * c + 4 * cc], lw[16 * c + 4 * cc + 1], lw[16 * c + 4 * cc + 2], lw[16 * c + 4 * cc + 3]);
                    }
                }
                fence_proxy_async_smem();
                mbar_arrive(p_full(w));
            }
            drain_o();                                        // the last (possibly partial) group
            // ---- epilogue: O / l -> fp16x2 -> global (one 128-byte row piece per thread and part) ----
            if ((int)w * BQ + row < it.q_rows) {
                const float inv = 1.0f / lsum;
                uint16_t* oh = p.Os + (size_t)(it.q_row0 + w * BQ + row) * p.ldo + it.head * 64;
                uint16_t* ol = oh + p.os_stride;
#pragma unroll
                for (int cc = 0; cc < 8; ++cc) {
                    uint32_t h4[4], l4[4];
#pragma unroll
                    for (int u = 0; u < 4; ++u) split_pair(oacc[8 * cc + 2 * u] * inv, oacc[8 * cc + 2 * u + 1] * inv, h4[u], l4[u]);
                    *reinterpret_cast<uint4*>(oh + 8 * cc) = make_uint4(h4[0], h4[1], h4[2], h4[3]);
                    *reinterpret_cast<uint4*>(ol + 8 * cc) = make_uint4(l4[0], l4[1], l4[2], l4[3]);
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, TMEM_COLS);
}

}  // namespace al
}  // namespace ffb
