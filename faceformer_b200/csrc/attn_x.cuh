// attn_x.cuh -- decoder attention cores on the 5th-generation tensor cores (tcgen05 + TMEM + TMA), sm_100a only.
//
//   O = softmax(Q K^T / 8) V   per (group, head), head dim 64, <= 256 keys per group, in two flavours:
//   CROSS  group = wireframe: queries are all prefix positions of all its sequences, keys / values its rows of the
//          cross-attention cache (transformer.py:247-251; memory_key_padding_mask realised as "only valid rows exist";
//          the cache is computed once per wireframe).  K / V stay in shared memory while consecutive
//          work items share (wireframe, head).
//   SELF   decoder self-attention over the prefix, NO causal mask (transformer.py:242-246, model_para.py:222-223):
//          a work item is a tile of floor(128 / P) whole sequences; S = Q K^T is computed for the whole tile and the
//          softmax keeps only the block diagonal (row r attends to the keys of its own sequence).
//
// Operands are fp16x2 splits (x = h + l), every product is  l*h + h*l + h*h  (3 MMAs, fp32 accumulate in TMEM):
//   Q, K   rows of fp16x2 buffers [2][rows][ld] (written by GEMM epilogues / split once per wireframe), K-major smem
//          tiles of 64-byte rows with SWIZZLE_64B (the layout gemm_tc.cuh uses), 32 contraction elements per tile
//   V      the row-major rows themselves ([keys][head dim]), loaded as 128-byte rows with SWIZZLE_128B and consumed
//          through an MN-major B descriptor (no transposed copy of V exists anywhere)
//   P      softmax weights, written by the softmax warps into shared memory as fp16x2 (A operand of O += P V)
//
// One persistent CTA per SM walks a contiguous range of work items.  Roles (384 threads):
//   warp 0      TMA producer (Q per item, K / V per (group, head))
//   warp 1      TMEM allocator + S issuer (one thread): S_j = Q_j K^T into the TMEM buffer of warpgroup j & 1
//   warps 2, 3  P V issuers of warpgroup 0 / 1 (one thread each): O_j += P_c V_c as the softmax warpgroup publishes chunk c.
//               Three independent issuers keep every hand-off a plain mbarrier wait (no polling loop between a softmax
//               arrive and the MMA it unlocks), so the tensor pipe, both softmax warpgroups and the TMA loads overlap.
//   warps 4-7   softmax warpgroup 0 (items 0, 2, 4, ... of the CTA), warps 8-11 softmax warpgroup 1 (items 1, 3, ...):
//               one query row per thread (TMEM lane = row): pass 1 row max over S, pass 2 p = 2^(s*log2e/8 - m + 12)
//               per 32-key chunk -> fp16x2 -> smem (double-buffered per warpgroup), finally O / l -> fp16x2 -> coalesced
//               global stores through a per-warp staging area.
// TMEM: two S buffers of 256 columns; O (64 columns) aliases columns [0,64) of its S buffer: the first P V product is
// issued only after the softmax has read S chunks 0 and 1.
#pragma once
#include "gemm_tc.cuh"
#include "attn_h.cuh"   // split_pair, ex2_approx

namespace ffb {
namespace ax {

constexpr int BQ = 128, KMAX = 256, KC = 32, NUM_THREADS = 384, MAX_GROUPS = 255;
constexpr int Q_TILE = BQ * 64;                  // bytes of one (part, k-chunk) Q tile: 128 rows x 64 B
constexpr int V_TILE = KC * 128;                 // one 32-key chunk of V: 32 rows x 128 B
constexpr int P_TILE = BQ * 64;                  // one (buffer, part) P chunk: 128 rows x 32 keys
constexpr int Q_BYTES = 4 * Q_TILE, K_BYTES = 4 * KMAX * 64, V_BYTES = 2 * (KMAX / KC) * V_TILE, P_BYTES = 2 * 4 * P_TILE;
constexpr int BAR_BYTES = 256, TOFF_BYTES = (MAX_GROUPS + 1) * 4;
constexpr int SMEM_BYTES = Q_BYTES + K_BYTES + V_BYTES + P_BYTES + BAR_BYTES + TOFF_BYTES + 1024;
static_assert(SMEM_BYTES <= 232448, "exceeds 227 KB of shared memory");
constexpr int TMEM_COLS = 512;

struct Params {
    int mode;                                  // 0 = CROSS, 1 = SELF
    // CROSS: queries of group i are rows [seq_off[i]*q_mul, seq_off[i+1]*q_mul); keys rows [row_off[i], +vlen[i])
    const int* seq_off; int q_mul; const int* row_off; const int* vlen;
    int n_groups;
    // SELF: n_seqs sequences of P rows each; a tile holds seqs_per_tile = floor(128 / P) whole sequences
    int P, seqs_per_tile, n_seqs;
    int n_heads;
    int q_col, k_col, v_col;                   // first column of head 0 in the Q / K / V rows
    int total_items;
    uint16_t* Os; long long os_stride; int ldo;   // fp16x2 output [2][rows][ldo]
    const int* stop;
};

struct Item {
    int q_row0, q_rows;                        // first query row, valid rows of the tile (<= 128)
    int k_row0, nk;                            // first key row, keys (<= 256)
    int v0;                                    // first value row
    int head, key;                             // key identifies the (group, head) whose K / V are in shared memory
    int P;                                     // SELF: sequence length (band width); 0 = CROSS (every key is valid)
};

__device__ __forceinline__ bool mbar_test(uint32_t bar, uint32_t parity) {       // non-blocking: has the phase with this parity completed?
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ uint32_t idesc_f16(int n, int b_mn) {     // kind::f16, fp16 A/B, fp32 D, K-major A, M = 128
    return (1u << 4) | ((uint32_t)b_mn << 16) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(BQ >> 4) << 24);
}
// MN-major B tile: rows = 32 keys of 128 bytes (64 head dims), SWIZZLE_128B; 8-key groups are 1024 B apart
// (stride byte offset; the leading byte offset is unused because the 64 head dims are exactly one swizzle atom wide).
// Verified on B200 against the transposed-cache path (tests/test_gpu_ops.py::test_attention_tcgen05_cross kind 6).
__device__ __forceinline__ uint64_t make_smem_desc_mn128(uint32_t saddr) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);
    d |= (uint64_t)(4096 >> 4) << 16;
    d |= (uint64_t)(1024 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;                         // SWIZZLE_128B
    return d;
}

// idx -> work item.  `hint` carries the group of the previous lookup (CROSS, items are visited in increasing order).
__device__ __forceinline__ void get_item(const Params& p, const int* toff, int idx, int& hint, Item& it) {
    const int H = p.n_heads;
    if (p.mode == 0) {
        int wf = hint;
        while (wf + 1 < p.n_groups && idx >= H * toff[wf + 1]) ++wf;
        hint = wf;
        const int tiles = toff[wf + 1] - toff[wf];
        const int rem = idx - H * toff[wf];
        const int head = rem / tiles, tile = rem - head * tiles;
        const long long q0 = (long long)p.seq_off[wf] * p.q_mul, q1 = (long long)p.seq_off[wf + 1] * p.q_mul;
        it.q_row0 = (int)(q0 + (long long)tile * BQ);
        it.q_rows = (int)min((long long)BQ, q1 - it.q_row0);
        it.k_row0 = p.row_off[wf]; it.nk = p.vlen[wf];
        it.v0 = it.k_row0;
        it.head = head; it.key = wf * H + head; it.P = 0;
    } else {
        const int tile = idx / H, head = idx - tile * H;
        const int s0 = tile * p.seqs_per_tile, ns = min(p.seqs_per_tile, p.n_seqs - s0);
        it.q_row0 = s0 * p.P; it.q_rows = ns * p.P;
        it.k_row0 = it.q_row0; it.nk = it.q_rows; it.v0 = it.q_row0;
        it.head = head; it.key = idx; it.P = p.P;
    }
}

__global__ void __launch_bounds__(NUM_THREADS, 1)
attn_x_kernel(const __grid_constant__ CUtensorMap mapQ, const __grid_constant__ CUtensorMap mapK,
              const __grid_constant__ CUtensorMap mapV, const Params p) {
    using namespace tc;
    FFB_PDL_SYNC();
    if (p.stop != nullptr && *p.stop != 0) return;

    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
    const uint32_t q_s = smem_base, k_s = q_s + Q_BYTES, v_s = k_s + K_BYTES, p_s = v_s + V_BYTES;
    const uint32_t bar_base = p_s + P_BYTES;
    uint8_t* p_gen = smem_gen + (p_s - smem_base);
    int* toff = reinterpret_cast<int*>(smem_gen + (bar_base - smem_base) + BAR_BYTES);
    // barriers (8 bytes each)
    const uint32_t q_full = bar_base, q_empty = bar_base + 8;
    auto kv_full = [&](uint32_t s) { return bar_base + 16 + 8 * s; };
    auto kv_empty = [&](uint32_t s) { return bar_base + 32 + 8 * s; };
    auto s_full = [&](uint32_t w) { return bar_base + 48 + 8 * w; };
    auto o_full = [&](uint32_t w) { return bar_base + 64 + 8 * w; };
    auto tmem_empty = [&](uint32_t w) { return bar_base + 80 + 8 * w; };
    auto p_full = [&](uint32_t w, uint32_t b) { return bar_base + 96 + 8 * (2 * w + b); };
    auto p_empty = [&](uint32_t w, uint32_t b) { return bar_base + 128 + 8 * (2 * w + b); };
    const uint32_t tmem_slot = bar_base + 160;
    uint32_t* tmem_slot_ptr = reinterpret_cast<uint32_t*>(smem_gen + (tmem_slot - smem_base));
    volatile int* done_cnt = reinterpret_cast<volatile int*>(smem_gen + (bar_base - smem_base) + 168);   // items finished per softmax warpgroup

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // K / V slots: CROSS keeps one (group, head) resident (<= 256 keys); SELF double-buffers per item (<= 128 keys)
    const uint32_t n_slots = p.mode == 0 ? 1u : 2u;
    const uint32_t k_tile = p.mode == 0 ? KMAX * 64 : (KMAX / 2) * 64;           // bytes of one (part, k-chunk) K tile
    const uint32_t k_slot = 4 * k_tile, v_chunks = p.mode == 0 ? KMAX / KC : KMAX / KC / 2, v_slot = 2 * v_chunks * V_TILE;

    if (p.mode == 0) {                                   // tiles per group -> exclusive prefix in shared memory
        for (int i = threadIdx.x; i < p.n_groups; i += NUM_THREADS) {
            const long long nq = ((long long)p.seq_off[i + 1] - p.seq_off[i]) * p.q_mul;
            toff[i + 1] = (int)((nq + BQ - 1) / BQ);
        }
    }
    if (threadIdx.x == 0) {
        mbar_init(q_full, 1); mbar_init(q_empty, 1);
        for (uint32_t s = 0; s < 2; ++s) {
            mbar_init(kv_full(s), 1); mbar_init(kv_empty(s), 1);
            mbar_init(s_full(s), 1); mbar_init(o_full(s), 1); mbar_init(tmem_empty(s), 128);
            mbar_init(p_full(s, 0), 128); mbar_init(p_full(s, 1), 128); mbar_init(p_empty(s, 0), 1); mbar_init(p_empty(s, 1), 1);
        }
        done_cnt[0] = 0; done_cnt[1] = 0;
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) tmem_alloc(tmem_slot, TMEM_COLS);
    __syncthreads();
    if (threadIdx.x == 0 && p.mode == 0) {
        int acc = 0; toff[0] = 0;
        for (int i = 1; i <= p.n_groups; ++i) { acc += toff[i]; toff[i] = acc; }
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot_ptr;

    const int item0 = (int)((long long)blockIdx.x * p.total_items / gridDim.x);
    const int n_items = (int)((long long)(blockIdx.x + 1) * p.total_items / gridDim.x) - item0;

    if (warp == 0) {
        if (lane == 0) {
            // ===== TMA producer =====
            int hint = 0, cur_key = -1; uint32_t n_kv = 0;
            for (int j = 0; j < n_items; ++j) {
                Item it; get_item(p, toff, item0 + j, hint, it);
                if (it.key != cur_key) {
                    cur_key = it.key;
                    const uint32_t slot = n_kv % n_slots;
                    // every item that read the previous content of this slot must have finished its P V products: the softmax
                    // warpgroups publish their completed-item counts after observing o_full
                    const int need0 = (n_slots == 1u) ? (j + 1) / 2 : ((j & 1) ? 0 : j / 2);
                    const int need1 = (n_slots == 1u) ? j / 2 : ((j & 1) ? j / 2 : 0);
                    while (done_cnt[0] < need0 || done_cnt[1] < need1) __nanosleep(64);
                    __threadfence_block();
                    fence_proxy_async_smem();
                    const int nb = (it.nk + KC - 1) / KC;
                    mbar_expect_tx(kv_full(slot), (uint32_t)nb * (4u * 2048u + 2u * V_TILE));
                    const uint32_t ks = k_s + slot * k_slot, vs = v_s + slot * v_slot;
                    for (int part = 0; part < 2; ++part)
                        for (int kch = 0; kch < 2; ++kch)
                            for (int b2 = 0; b2 < nb; ++b2)
                                tma_load_3d(ks + (part * 2 + kch) * k_tile + b2 * 2048, &mapK, kv_full(slot),
                                            p.k_col + it.head * 64 + kch * 32, it.k_row0 + b2 * KC, part);
                    for (int part = 0; part < 2; ++part)
                        for (int b2 = 0; b2 < nb; ++b2) {
                            const uint32_t dst = vs + (part * v_chunks + b2) * V_TILE;
                            tma_load_3d(dst, &mapV, kv_full(slot), p.v_col + it.head * 64, it.v0 + b2 * KC, part);
                        }
                    ++n_kv;
                }
                while (!mbar_test(q_empty, ((uint32_t)j & 1u) ^ 1u)) __nanosleep(32);
                mbar_expect_tx(q_full, Q_BYTES);
                for (int part = 0; part < 2; ++part)
                    for (int kch = 0; kch < 2; ++kch)
                        tma_load_3d(q_s + (part * 2 + kch) * Q_TILE, &mapQ, q_full, p.q_col + it.head * 64 + kch * 32, it.q_row0, part);
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            // ===== S issuer: S_j = Q_j K^T into the TMEM buffer of warpgroup j & 1 =====
            constexpr int pa[3] = {1, 0, 0}, pb[3] = {0, 1, 0};            // (lo,hi), (hi,lo), (hi,hi): small products first
            int hint = 0, cur_key = -1; uint32_t n_kv = 0, slot = 0;
            for (int j = 0; j < n_items; ++j) {
                Item it; get_item(p, toff, item0 + j, hint, it);
                const uint32_t w = (uint32_t)j & 1u;
                if (it.key != cur_key) {
                    cur_key = it.key;
                    slot = n_kv % n_slots;
                    mbar_wait(kv_full(slot), (n_kv / n_slots) & 1u);
                    ++n_kv;
                }
                mbar_wait(q_full, (uint32_t)j & 1u);
                mbar_wait(tmem_empty(w), (((uint32_t)j >> 1) & 1u) ^ 1u);  // S / O of this warpgroup's previous item have been read out
                tc_fence_after();
                const uint32_t d_s = tmem_base + w * 256u;
                const uint32_t idesc_s = idesc_f16((it.nk + 15) & ~15, 0);
                const uint32_t ks = k_s + slot * k_slot;
                uint32_t acc = 0;
#pragma unroll
                for (int kch = 0; kch < 2; ++kch)
#pragma unroll
                    for (int s2 = 0; s2 < 2; ++s2)
#pragma unroll
                        for (int q = 0; q < 3; ++q) {
                            const uint64_t da = make_smem_desc(q_s + (pa[q] * 2 + kch) * Q_TILE) + (uint64_t)(2 * s2);
                            const uint64_t db = make_smem_desc(ks + (pb[q] * 2 + kch) * k_tile) + (uint64_t)(2 * s2);
                            umma_bf16(d_s, da, db, idesc_s, acc);
                            acc = 1;
                        }
                umma_commit(q_empty);
                umma_commit(s_full(w));
            }
        }
    } else if (warp < 4) {
        if (lane == 0) {
            // ===== P V issuer of warpgroup w: O_j += P_c V_c as the softmax warpgroup publishes the chunks =====
            constexpr int pa[3] = {1, 0, 0}, pb[3] = {0, 1, 0};
            const uint32_t w = (uint32_t)warp - 2u;
            const uint32_t idesc_o = idesc_f16(64, 1);
            const uint32_t d_o = tmem_base + w * 256u;                     // aliases S columns [0,64) of this warpgroup
            int hint = 0, cur_key = -1; uint32_t n_kv = 0, slot = 0, c = 0;
            for (int j = 0; j < n_items; ++j) {
                Item it; get_item(p, toff, item0 + j, hint, it);
                if (it.key != cur_key) { cur_key = it.key; slot = n_kv % n_slots; ++n_kv; }
                if (((uint32_t)j & 1u) != w) continue;
                const int nb = (it.nk + KC - 1) / KC;
                const uint32_t vs = v_s + slot * v_slot;
                uint32_t acc = 0;
                for (int b2 = 0; b2 < nb; ++b2, ++c) {
                    const uint32_t buf = c & 1u;
                    mbar_wait(p_full(w, buf), (c >> 1) & 1u);
                    // O aliases S columns [0,64): the first product must not start before S chunk 1 has been read (chunk 1 published)
                    if (b2 == 0 && nb > 1) mbar_wait(p_full(w, buf ^ 1u), ((c + 1) >> 1) & 1u);
                    tc_fence_after();
#pragma unroll
                    for (int s2 = 0; s2 < 2; ++s2)
#pragma unroll
                        for (int q = 0; q < 3; ++q) {
                            const uint64_t da = make_smem_desc(p_s + ((w * 2 + buf) * 2 + pa[q]) * P_TILE) + (uint64_t)(2 * s2);
                            const uint32_t vb = vs + (pb[q] * v_chunks + b2) * V_TILE;
                            const uint64_t db = make_smem_desc_mn128(vb + s2 * 2048u);
                            umma_bf16(d_o, da, db, idesc_o, acc);
                            acc = 1;
                        }
                    umma_commit(p_empty(w, buf));
                }
                umma_commit(o_full(w));
            }
        }
    } else {
        // ===== softmax + epilogue warpgroups: thread = query row =====
        const uint32_t w = (uint32_t)(warp - 4) >> 2;        // warpgroup 0 / 1
        const int q = warp & 3;                               // TMEM lane quarter this warp may access
        const int row = q * 32 + lane;
        const uint32_t t_s = tmem_base + ((uint32_t)(q * 32) << 16) + w * 256u;
        constexpr float kScale = 0.125f * 1.4426950408889634f;
        uint8_t* pw = p_gen + w * 4 * P_TILE;                 // this warpgroup's P tiles [buf][part]
        // per-warp output staging: four 2 KB pieces that coincide with this warp's own rows of its warpgroup's four P tiles
        auto piece = [&](int k) { return pw + k * P_TILE + q * 2048; };
        int hint = 0; uint32_t n_pc = 0;
        for (int j = (int)w, k = 0; j < n_items; j += 2, ++k) {
            Item it; get_item(p, toff, item0 + j, hint, it);
            const int nb = (it.nk + KC - 1) / KC;
            // keys this row attends to, and the union over the warp's rows
            int klo = 0, khi = it.nk, wlo = 0, whi = it.nk;
            if (it.P > 0) {
                const bool valid = row < it.q_rows;
                klo = valid ? (row / it.P) * it.P : 0; khi = valid ? klo + it.P : 0;
                const int r0 = q * 32, r1 = min(q * 32 + 31, it.q_rows - 1);
                wlo = (r0 / it.P) * it.P; whi = (r1 >= r0) ? (r1 / it.P) * it.P + it.P : 0;
                if (r1 < r0) wlo = 0;
            }
            mbar_wait(s_full(w), (uint32_t)k & 1u);
            tc_fence_after();
            // valid-key bit mask of this row within chunk b (CROSS: the same for every row; SELF: the row's own sequence)
            auto chunk_mask = [&](int b) -> uint32_t {
                const int lo = max(klo - b * KC, 0), hi = min(khi - b * KC, KC);
                if (hi <= lo) return 0u;
                return (hi - lo >= 32) ? 0xffffffffu : (((1u << (hi - lo)) - 1u) << lo);
            };
            // ---- pass 1: row maximum over the valid keys ----
            float mx = -INFINITY;
            for (int b = 0; b < nb; ++b) {
                if (b * KC >= whi || b * KC + KC <= wlo) continue;           // warp-uniform: no row of this warp attends to this chunk
                uint32_t v[32];
                tmem_ld32(t_s + b * KC, v);
                tmem_ld_wait();
                const uint32_t M = chunk_mask(b);
                if (M == 0xffffffffu) {                                     // whole chunk valid for this row
                    float m0 = mx, m1 = -INFINITY;
#pragma unroll
                    for (int i = 0; i < 32; i += 4) {
                        m0 = fmaxf(m0, fmaxf(__uint_as_float(v[i]), __uint_as_float(v[i + 1])));
                        m1 = fmaxf(m1, fmaxf(__uint_as_float(v[i + 2]), __uint_as_float(v[i + 3])));
                    }
                    mx = fmaxf(m0, m1);
                } else if (M != 0u) {
#pragma unroll
                    for (int i = 0; i < 32; ++i) mx = fmaxf(mx, (M & (1u << i)) ? __uint_as_float(v[i]) : -INFINITY);
                }
            }
            const float bias = 12.0f - mx * kScale;           // p' = 2^(s*kScale - m' + 12) = 4096 * exp((s - m)/8)
            float lsum = 0.f;
            // ---- pass 2: probabilities per 32-key chunk -> fp16x2 -> shared memory (A operand of O += P V) ----
            for (int b = 0; b < nb; ++b, ++n_pc) {
                const uint32_t buf = n_pc & 1u;
                uint8_t* ph = pw + (buf * 2 + 0) * P_TILE + row * 64;
                uint8_t* pl = pw + (buf * 2 + 1) * P_TILE + row * 64;
                const int rsw = (row >> 1) & 3;                           // 16-byte chunk cc lives at cc ^ ((row >> 1) & 3) (SWIZZLE_64B)
                // each case waits for the buffer (the MMAs that read it two chunks ago have retired) and stores its own words, so the
                // all-zero cases do not drag 32 register moves through the hot path
                auto publish = [&](const uint32_t (&hw)[16], const uint32_t (&lw)[16]) {
                    mbar_wait(p_empty(w, buf), ((n_pc >> 1) & 1u) ^ 1u);
#pragma unroll
                    for (int cc = 0; cc < 4; ++cc) {
                        const int sw = (cc ^ rsw) << 4;
                        *reinterpret_cast<uint4*>(ph + sw) = make_uint4(hw[4 * cc], hw[4 * cc + 1], hw[4 * cc + 2], hw[4 * cc + 3]);
                        *reinterpret_cast<uint4*>(pl + sw) = make_uint4(lw[4 * cc], lw[4 * cc + 1], lw[4 * cc + 2], lw[4 * cc + 3]);
                    }
                };
                auto publish_zero = [&]() {
                    mbar_wait(p_empty(w, buf), ((n_pc >> 1) & 1u) ^ 1u);
                    const uint4 z = make_uint4(0u, 0u, 0u, 0u);
#pragma unroll
                    for (int cc = 0; cc < 4; ++cc) { *reinterpret_cast<uint4*>(ph + (cc << 4)) = z; *reinterpret_cast<uint4*>(pl + (cc << 4)) = z; }
                };
                if (b * KC >= whi || b * KC + KC <= wlo) {
                    publish_zero();
                } else {
                    uint32_t v[32];
                    tmem_ld32(t_s + b * KC, v);
                    tmem_ld_wait();
                    const uint32_t M = chunk_mask(b);
                    if (M == 0xffffffffu) {                                 // whole chunk valid: no selects
                        uint32_t hw[16], lw[16];
                        float ls0 = 0.f, ls1 = 0.f;
#pragma unroll
                        for (int i = 0; i < 16; ++i) {
                            const float p0 = ex2_approx(fmaf(__uint_as_float(v[2 * i]), kScale, bias));
                            const float p1 = ex2_approx(fmaf(__uint_as_float(v[2 * i + 1]), kScale, bias));
                            ls0 += p0; ls1 += p1;
                            split_pair(p0, p1, hw[i], lw[i]);
                        }
                        lsum += ls0 + ls1;
                        publish(hw, lw);
                    } else if (M != 0u) {
                        uint32_t hw[16], lw[16];
                        float ls0 = 0.f, ls1 = 0.f;
#pragma unroll
                        for (int i = 0; i < 16; ++i) {
                            float p0 = ex2_approx(fmaf(__uint_as_float(v[2 * i]), kScale, bias));
                            float p1 = ex2_approx(fmaf(__uint_as_float(v[2 * i + 1]), kScale, bias));
                            p0 = (M & (1u << (2 * i))) ? p0 : 0.f;
                            p1 = (M & (2u << (2 * i))) ? p1 : 0.f;
                            ls0 += p0; ls1 += p1;
                            split_pair(p0, p1, hw[i], lw[i]);
                        }
                        lsum += ls0 + ls1;
                        publish(hw, lw);
                    } else {
                        publish_zero();
                    }
                }
                fence_proxy_async_smem();
                tc_fence_before();                                        // this thread's S reads precede the O writes the arrive unlocks
                mbar_arrive(p_full(w, buf));
            }
            // ---- epilogue: O / l -> fp16x2 -> global ----
            mbar_wait(o_full(w), (uint32_t)k & 1u);
            tc_fence_after();
            uint32_t o0[32], o1[32];
            tmem_ld32(t_s, o0);
            tmem_ld32(t_s + 32, o1);
            tmem_ld_wait();
            tc_fence_before();
            mbar_arrive(tmem_empty(w));
            if ((warp & 3) == 0 && lane == 0) { __threadfence_block(); done_cnt[w] = k + 1; }   // this warpgroup's P V products of item j have retired
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
                    if (q * 32 + r < it.q_rows)
                        *reinterpret_cast<uint4*>(p.Os + (size_t)part * p.os_stride + (size_t)(it.q_row0 + q * 32 + r) * p.ldo + it.head * 64 + ch * 8) = val;
                }
            __syncwarp();
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, TMEM_COLS);
}

}  // namespace ax
}  // namespace ffb
