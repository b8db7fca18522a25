// gemm_tc2.cuh -- the fp16x2 split-precision GEMM of gemm_tc.cuh on CTA PAIRS (tcgen05 cta_group::2), sm_100a only.
//
// Why: profiles/tune_gemm.py shows the single-CTA kernel bound by L2 -> SM operand traffic, not by the tensor pipe or by
// its epilogue: with the TMA stores suppressed ("dry" runs) it reaches the mainloop-only rate, and every byte of output
// written through L2 costs mainloop throughput 1:1.  Per 128 x 256 x 32 MMA block a single CTA pulls 16 KB of A and 32 KB of
// W (both splits) out of L2.  A CTA pair computes a 256 x 256 tile with ONE tcgen05.mma.cta_group::2 per k-step: each CTA
// loads its own 128 rows of A and only HALF of the W tile (128 of the 256 output columns), the tensor cores of the two SMs
// read both halves from each other's shared memory: 32 KB instead of 48 KB per CTA and k-block (-33 % L2 -> SM traffic,
// -33 % shared-memory fill).
//
// STATUS (round 1): numerically green (tests/test_gpu_tc.py, fmt 23) but SLOWER than the single-CTA kernel -- 273-305 vs 360-395
// TFLOP/s in profiles/tune_gemm.py --random, tensor pipe 39 % active in ncu (vs 60-65 %), with neither the MMA thread nor the
// producers waiting on barriers -- so it is NOT the default (FFB_OPT_GEMM_VARIANT = 3 selects it).  Kept as the starting point
// for the next round: 128-byte-swizzle operand tiles with BK = 64 (half as many, twice as large MMAs and TMA boxes) is the first
// thing to try.
//
// Same arithmetic as gemm_kernel<2>: products lo*hi + hi*lo + hi*hi per k-step, DRAIN_KB = 4 k-blocks accumulated in TMEM
// (24-MMA chains), then drained by the epilogue warps into fp32 register accumulators with round-to-nearest adds.
//
// Cluster of 2 CTAs (rank 0 = leader), 384 threads each:
//   warp 0      TMA producer of THIS CTA's operand halves; every load signals the LEADER's full barrier
//               (cp.async.bulk.tensor ... .cta_group::2); the leader's barrier expects the bytes of both CTAs
//   warp 1      leader only: MMA issuer (one thread); tcgen05.commit ... multicast::cluster releases the smem stage / publishes
//               the accumulator in BOTH CTAs
//   warp 2      TMEM allocator (cta_group::2; 512 columns = two ping-pong accumulators of 128 lanes x 256 columns per CTA)
//   warps 4-11  epilogue of this CTA's 128 rows: tcgen05.ld -> registers -> scale / bias / ReLU -> staging -> TMA store,
//               reduce-add or fp16x2 split store; drained accumulators are handed back with a remote arrive on the leader's barrier
#pragma once
#include "gemm_tc.cuh"

namespace ffb {
namespace tc2 {

using tc::BK;
constexpr int BM = 128, BN = 256;                    // per CTA: 128 rows x 256 columns of the 256 x 256 pair tile
constexpr int A_TILE_BYTES = BM * BK * 2;            // 8 KB per split
constexpr int B_HALF_BYTES = (BN / 2) * BK * 2;      // 8 KB per split: this CTA's 128 of the 256 W rows
#ifndef FFB_TC2_STAGES
#define FFB_TC2_STAGES 4
#endif
#ifndef FFB_TC2_EPI_BUFS
#define FFB_TC2_EPI_BUFS 2
#endif
constexpr int STAGES = FFB_TC2_STAGES, DRAIN_KB = 4;
constexpr int STAGE_BYTES = 2 * (A_TILE_BYTES + B_HALF_BYTES);        // 32 KB
constexpr int EPI_WARP0 = 4, EPI_WARPS = 8, EPI_BUFS = FFB_TC2_EPI_BUFS, NUM_THREADS = 384;
constexpr int EPI_BYTES = EPI_BUFS * EPI_WARPS * 4096;                // 64 KB
constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + EPI_BYTES + 1024 + 256;
static_assert(SMEM_BYTES <= 232448, "exceeds 227 KB of shared memory");
constexpr int TMEM_COLS = 512;
// kind::f16 instruction descriptor: D = f32, A / B = f16, K-major, N = 256, M = 256 (128 rows per CTA of the pair)
constexpr uint32_t IDESC = (1u << 4) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);

__device__ __forceinline__ uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ uint32_t mapa(uint32_t addr, uint32_t rank) {       // shared::cta address -> shared::cluster address in CTA `rank`
    uint32_t r; asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank)); return r;
}
__device__ __forceinline__ void cluster_sync() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// Relaxed remote arrive: no cluster-scope release fence (which would wait for the thread's outstanding bulk stores).  Enough for
// handing a drained TMEM accumulator back: the tcgen05.ld results are ordered by tcgen05.fence::before_thread_sync on this side
// and tcgen05.fence::after_thread_sync on the MMA side.
__device__ __forceinline__ void mbar_arrive_cluster_relaxed(uint32_t cluster_addr) {
    asm volatile("mbarrier.arrive.relaxed.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void tma_load_3d_2sm(uint32_t dst, const CUtensorMap* map, uint32_t leader_bar, int c0, int c1, int c2) {
    asm volatile("cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                 ::"r"(dst), "l"(map), "r"(leader_bar), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ void tmem_alloc_2sm(uint32_t dst_smem, uint32_t cols) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_2sm(uint32_t taddr, uint32_t cols) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(cols) : "memory");
}
__device__ __forceinline__ void umma_2sm(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit_2sm(uint32_t bar) {                  // arrives on `bar` in BOTH CTAs of the pair
    const uint16_t mask = 3;
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(bar), "h"(mask) : "memory");
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(NUM_THREADS, 1)
gemm2_kernel(const __grid_constant__ CUtensorMap mapA0, const __grid_constant__ CUtensorMap mapA1,
             const __grid_constant__ CUtensorMap mapW, const __grid_constant__ CUtensorMap mapC, const tc::Params p) {
    using namespace tc;
    if (p.stop != nullptr && *p.stop != 0) return;          // uniform over the grid: both CTAs of a pair leave together

    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t epi_base = smem_base + STAGES * STAGE_BYTES;
    const uint32_t bar_base = epi_base + EPI_BYTES;
    auto full_bar = [&](int s) { return bar_base + 8u * s; };
    auto empty_bar = [&](int s) { return bar_base + 8u * (STAGES + s); };
    auto tfull_bar = [&](int b) { return bar_base + 8u * (2 * STAGES + b); };
    auto tempty_bar = [&](int b) { return bar_base + 8u * (2 * STAGES + 2 + b); };
    const uint32_t tmem_slot = bar_base + 8u * (2 * STAGES + 4);
    uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
    uint32_t* tmem_slot_ptr = reinterpret_cast<uint32_t*>(smem_gen + (tmem_slot - smem_base));

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const bool leader = (rank == 0);
    const int pair = blockIdx.x >> 1, n_pairs = gridDim.x >> 1;
    const int m_tiles = (p.M + 2 * BM - 1) / (2 * BM), n_tiles = p.N / BN, k_chunks = p.K / BK;
    const int num_tiles = m_tiles * n_tiles;

    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; ++s) { mbar_init(full_bar(s), 2); mbar_init(empty_bar(s), 1); }        // full: both producers arrive
        for (int b = 0; b < 2; ++b) { mbar_init(tfull_bar(b), 1); mbar_init(tempty_bar(b), 2 * EPI_WARPS); }  // tempty: both CTAs' epilogue warps
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 2) tmem_alloc_2sm(tmem_slot, TMEM_COLS);
    tc_fence_before();
    __syncthreads();
    cluster_sync();                                          // the peer's barriers are initialised before anyone signals them
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot_ptr;

    if (warp < EPI_WARP0) {
        asm volatile("setmaxnreg.dec.sync.aligned.u32 40;");
        if (warp == 0 && lane == 0) {
            // ===== TMA producer (both CTAs) =====
            int stage = 0; uint32_t phase = 0;
            for (int t = pair; t < num_tiles; t += n_pairs) {
                const int mt = t / n_tiles, nt = t % n_tiles;
                const CUtensorMap* mapA = (nt >= p.n_switch) ? &mapA1 : &mapA0;
                const int row_a = mt * 2 * BM + (int)rank * BM, row_w = nt * BN + (int)rank * (BN / 2) + p.n_off;
                for (int kc = 0; kc < k_chunks; ++kc) {
                    mbar_wait(empty_bar(stage), phase ^ 1u);
                    const uint32_t lbar = mapa(full_bar(stage), 0);
                    if (leader) mbar_expect_tx(full_bar(stage), 2 * STAGE_BYTES);      // bytes of both CTAs land on the leader's barrier
                    else mbar_arrive_cluster(lbar);
                    const uint32_t sa = smem_base + stage * STAGE_BYTES;
                    const uint32_t sb = sa + 2 * A_TILE_BYTES;
#pragma unroll
                    for (int s = 0; s < 2; ++s) tma_load_3d_2sm(sa + s * A_TILE_BYTES, mapA, lbar, kc * BK, row_a, s);
#pragma unroll
                    for (int s = 0; s < 2; ++s) tma_load_3d_2sm(sb + s * B_HALF_BYTES, &mapW, lbar, kc * BK, row_w, s);
                    if (++stage == STAGES) { stage = 0; phase ^= 1u; }
                }
            }
        } else if (warp == 1 && lane == 0 && leader) {
            // ===== MMA issuer (leader CTA) =====
            int stage = 0; uint32_t phase = 0; uint32_t c = 0;
            for (int t = pair; t < num_tiles; t += n_pairs) {
                for (int kc = 0; kc < k_chunks; ++kc) {
                    const uint32_t buf = c & 1u;
                    const bool first_in_drain = (kc % DRAIN_KB) == 0;
                    const bool last_in_drain = ((kc % DRAIN_KB) == DRAIN_KB - 1) || (kc == k_chunks - 1);
                    if (first_in_drain) mbar_wait(tempty_bar(buf), ((c >> 1) & 1u) ^ 1u);
                    mbar_wait(full_bar(stage), phase);
                    tc_fence_after();
                    const uint32_t sa = smem_base + stage * STAGE_BYTES;
                    const uint32_t sb = sa + 2 * A_TILE_BYTES;
                    const uint32_t d = tmem_base + buf * BN;
                    constexpr int pa2[3] = {1, 0, 0}, pb2[3] = {0, 1, 0};   // correction products first, hi*hi last
                    uint32_t acc = first_in_drain ? 0u : 1u;
#pragma unroll
                    for (int q = 0; q < 3; ++q) {
                        const uint64_t da = make_smem_desc(sa + pa2[q] * A_TILE_BYTES);
                        const uint64_t db = make_smem_desc(sb + pb2[q] * B_HALF_BYTES);
#pragma unroll
                        for (int k = 0; k < BK / 16; ++k) {
                            umma_2sm(d, da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), IDESC, acc);
                            acc = 1;
                        }
                    }
                    umma_commit_2sm(empty_bar(stage));
                    if (last_in_drain) { umma_commit_2sm(tfull_bar(buf)); ++c; }
                    if (++stage == STAGES) { stage = 0; phase ^= 1u; }
                }
            }
        }
    } else {
        // ===== epilogue warps (both CTAs): this CTA's 128 rows x 256 columns =====
        asm volatile("setmaxnreg.inc.sync.aligned.u32 232;");
        const int e = warp - EPI_WARP0;
        const int q = warp & 3;
        const int hcol = e >> 2;
        float* stg0 = reinterpret_cast<float*>(smem_gen + (epi_base - smem_base)) + e * 1024 * EPI_BUFS;
        uint32_t n_blk = 0;
        float acc[128];
        uint32_t c = 0;
        const bool to_split = (p.C == nullptr) && (p.Cs != nullptr);
        for (int t = pair; t < num_tiles; t += n_pairs) {
            const int mt = t / n_tiles, nt = t % n_tiles;
            const int n_drains = (k_chunks + DRAIN_KB - 1) / DRAIN_KB;
            for (int dr = 0; dr < n_drains; ++dr, ++c) {
                const uint32_t buf = c & 1u;
                mbar_wait(tfull_bar(buf), (c >> 1) & 1u);
                tc_fence_after();
                const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + buf * BN + hcol * 128;
#pragma unroll
                for (int j = 0; j < 2; ++j) {
                    uint32_t v0[32], v1[32];
                    tmem_ld32(taddr + j * 64, v0);
                    tmem_ld32(taddr + j * 64 + 32, v1);
                    tmem_ld_wait();
                    if (dr == 0) {
#pragma unroll
                        for (int i = 0; i < 32; ++i) { acc[j * 64 + i] = __uint_as_float(v0[i]); acc[j * 64 + 32 + i] = __uint_as_float(v1[i]); }
                    } else {
#pragma unroll
                        for (int i = 0; i < 32; ++i) { acc[j * 64 + i] += __uint_as_float(v0[i]); acc[j * 64 + 32 + i] += __uint_as_float(v1[i]); }
                    }
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive_cluster_relaxed(mapa(tempty_bar(buf), 0));      // the leader's MMA thread owns the accumulators
            }
            const int col0 = nt * BN + hcol * 128 + p.n_off;
            const int row0 = mt * 2 * BM + (int)rank * BM + q * 32;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const uint32_t sbuf = (EPI_BUFS == 2) ? (n_blk & 1u) : 0u;
                float* stg = stg0 + sbuf * 1024;
                const uint32_t stg_s = epi_base + ((uint32_t)e * EPI_BUFS + sbuf) * 4096u;
                if (lane == 0) tma_store_wait_read<EPI_BUFS - 1>();
                __syncwarp();
                ++n_blk;
                if (to_split) {
                    float amax = 0.f;
#pragma unroll
                    for (int cc = 0; cc < 4; ++cc) {
                        float xv[8];
                        if (p.bias) {
                            const float4 b0 = __ldg(reinterpret_cast<const float4*>(p.bias + col0 + j * 32 + 8 * cc));
                            const float4 b1 = __ldg(reinterpret_cast<const float4*>(p.bias + col0 + j * 32 + 8 * cc + 4));
                            const float bv[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
                            for (int u = 0; u < 8; ++u) xv[u] = fmaf(acc[j * 32 + 8 * cc + u], p.out_scale, bv[u]);
                        } else {
#pragma unroll
                            for (int u = 0; u < 8; ++u) xv[u] = acc[j * 32 + 8 * cc + u] * p.out_scale;
                        }
                        if (p.relu) {
#pragma unroll
                            for (int u = 0; u < 8; ++u) xv[u] = fmaxf(xv[u], 0.f);
                        }
                        uint32_t hw[4], lw[4];
#pragma unroll
                        for (int u = 0; u < 4; ++u) {
                            const float x0 = xv[2 * u], x1 = xv[2 * u + 1];
                            amax = fmaxf(amax, fmaxf(fabsf(x0), fabsf(x1)));
                            const __half2 h = __floats2half2_rn(x0, x1);
                            const float2 hf = __half22float2(h);
                            const __half2 l = __floats2half2_rn(x0 - hf.x, x1 - hf.y);
                            hw[u] = *reinterpret_cast<const uint32_t*>(&h);
                            lw[u] = *reinterpret_cast<const uint32_t*>(&l);
                        }
                        uint8_t* base = reinterpret_cast<uint8_t*>(stg) + lane * 64 + ((cc ^ ((lane >> 1) & 3)) << 4);
                        *reinterpret_cast<uint4*>(base) = make_uint4(hw[0], hw[1], hw[2], hw[3]);
                        *reinterpret_cast<uint4*>(base + 2048) = make_uint4(lw[0], lw[1], lw[2], lw[3]);
                    }
                    if (!(amax <= 65504.f) && p.overflow) *p.overflow = 1;
                    fence_proxy_async_smem();
                    __syncwarp();
                    if (lane == 0 && !p.dry_store) {
                        tma_store_3d(&mapC, stg_s, col0 + j * 32, row0, 0);
                        tma_store_3d(&mapC, stg_s + 2048u, col0 + j * 32, row0, 1);
                        tma_store_commit();
                    }
                } else {
#pragma unroll
                    for (int cc = 0; cc < 8; ++cc) {
                        float4 v = make_float4(acc[j * 32 + 4 * cc] * p.out_scale, acc[j * 32 + 4 * cc + 1] * p.out_scale,
                                               acc[j * 32 + 4 * cc + 2] * p.out_scale, acc[j * 32 + 4 * cc + 3] * p.out_scale);
                        if (p.bias) {
                            const float4 b = __ldg(reinterpret_cast<const float4*>(p.bias + col0 + j * 32 + 4 * cc));
                            v.x += b.x; v.y += b.y; v.z += b.z; v.w += b.w;
                        }
                        if (p.relu) { v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f); }
                        *reinterpret_cast<float4*>(stg + lane * 32 + ((cc ^ (lane & 7)) << 2)) = v;
                    }
                    fence_proxy_async_smem();
                    __syncwarp();
                    if (lane == 0 && !p.dry_store) {
                        if (p.R) tma_reduce_add_2d(&mapC, stg_s, col0 + j * 32, row0);
                        else tma_store_2d(&mapC, stg_s, col0 + j * 32, row0);
                        tma_store_commit();
                    }
                }
            }
        }
        if (lane == 0) tma_store_wait_all();
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync();                                          // the peer may still be signalling this CTA's barriers / reading its smem
    if (warp == 2) tmem_dealloc_2sm(tmem_base, TMEM_COLS);
}

}  // namespace tc2
}  // namespace ffb
