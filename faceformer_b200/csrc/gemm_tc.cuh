// gemm_tc.cuh -- split-precision tensor-core GEMM for the decode-step linear layers (sm_100a only).
//
//   C[M,N] = epilogue( A[M,K] . W[N,K]^T )      A, W given as 16-bit splits of fp32 values (template NS):
//       NS = 3  bf16x3:  x = x0 + x1 + x2 (exact to 2^-26), 6 MMAs  A0W0 + A0W1 + A1W0 + A1W1 + A0W2 + A2W0
//       NS = 2  fp16x2:  x = h + l       (exact to 2^-22), 3 MMAs  hh + hl + lh; half the tensor work and 2/3 of the
//               operand bytes, but fp16 range: weights are pre-scaled by a power of two per matrix (undone in the
//               epilogue, exact) and producers raise an overflow flag if |activation| > 65504 (the host then re-runs
//               the decode in bf16x3).
//
// Why split precision: the parity contract (token-exact, logits within 1e-4 of the reference's fp32 CPU path)
// needs fp32-class products; one bf16 pass is ~6000x over budget.  Six bf16 MMAs
//   A0W0 + A0W1 + A1W0 + A1W1 + A0W2 + A2W0       (dropped terms <= 2^-26 relative)
// reproduce the fp32 product to better than fp32 rounding, at 1/6 of the bf16 tensor rate (~230 TFLOP/s of
// fp32-equivalent work at the measured 1382 TFLOP/s) instead of the ~37 TFLOP/s of the SIMT FFMA kernel.
//
// Accumulation: the tensor core adds into its fp32 accumulator with truncation (measured: a small negative
// bias on all-positive data), so chains are kept short: DRAIN_KB k-blocks (BK = 32 each) are accumulated in
// TMEM -- per k-block the five small correction products first, the dominant A0W0 product last -- then the
// epilogue warps drain the accumulator and add it into fp32 REGISTER accumulators with round-to-nearest
// FADDs.  DRAIN_KB = 2 balances the TMEM read port (a 128x256 fp32 drain costs ~2k cycles, measured) against
// the MMA time of the k-blocks it covers; chains stay 24 MMAs long independent of K.
//
// Epilogue: NS = 2 leaves room in shared memory for a per-warp 32x32 transposition buffer, so outputs (and the
// residual read) go out as fully coalesced 128-byte rows; NS = 3 (3 x 72 KB of operand stages) writes one row per
// thread directly (un-coalesced; measured 138 vs 350 TFLOP/s mainloop-only, profiles/tune_gemm.py).
//
// Structure (one CTA per SM, persistent over output tiles of 128 x 256):
//   warp 0      TMA producer: cp.async.bulk.tensor (3-D maps: k, row, split) into a 3/4-stage smem ring
//   warp 1      MMA issuer: one elected thread issues tcgen05.mma.cta_group::1.kind::f16 (M128 N256 K16)
//   warp 2      TMEM allocator (512 columns = two 128x256 fp32 accumulators, ping-pong per k-block)
//   warps 4-11  epilogue: tcgen05.ld 32x32b -> register accumulate -> bias / ReLU / residual -> global
// Pipelines: full/empty mbarriers (TMA <-> MMA) and tfull/tempty mbarriers (MMA <-> epilogue).
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "kernels.cuh"   // split3_bf16

namespace ffb {
namespace tc {

constexpr int BM = 128, BN = 256, BK = 32;          // BK 16-bit elements = 64-byte rows -> SWIZZLE_64B
constexpr int A_TILE_BYTES = BM * BK * 2;           //  8 KB
constexpr int B_TILE_BYTES = BN * BK * 2;           // 16 KB
constexpr int NUM_THREADS = 384;
constexpr int EPI_WARP0 = 4, EPI_WARPS = 8;
constexpr int TMEM_COLS = 512;

// V: pipeline variant of the fp16x2 kernel.  0 = 4 operand stages + one 4 KB epilogue staging buffer per warp;
//    1 = 3 operand stages + two staging buffers per warp (the next 32x32 block is formatted while TMA still reads the previous one)
// BNT: output-tile width.  256 (default) is the throughput shape; 64 is the "skinny" shape for small M (decode steps of one wireframe,
//      seq2seq): a 128 x 256 tile costs the tensor pipe 128 rows of work however few rows exist, and with M <= 128 only N / 256 CTAs would
//      run; 64-column tiles put 4x as many SMs on the weight stream and cut the per-tile MMA time 4x (fp16x2 only).
template <int NS, int V = 0, int BNT = BN> struct Cfg {
    static_assert(NS == 2 || NS == 3, "NS: 2 = fp16x2, 3 = bf16x3");
    static_assert(V == 0 || (V == 1 && NS == 2), "variant 1 exists for fp16x2 only");
    static_assert(BNT == 256 || (BNT == 64 && NS == 2), "tile width: 256, or 64 for fp16x2");
    static constexpr int B_TILE = BNT * BK * 2;
    static constexpr int STAGES = (BNT == 64) ? 6 : (NS == 2) ? (V == 1 ? 3 : 4) : 3;
    static constexpr int EPI_BUFS = (V == 1) ? 2 : 1;
#ifndef FFB_DRAIN_KB
#define FFB_DRAIN_KB 4
#endif
    static constexpr int DRAIN_KB = (NS == 2) ? FFB_DRAIN_KB : 2;  // k-blocks accumulated in TMEM per drain (24 MMAs either way by default)
    static constexpr int NPROD = (NS == 2) ? 3 : 6;
    static constexpr int STAGE_BYTES = NS * (A_TILE_BYTES + B_TILE);                  // 48 KB / 72 KB (24 KB for the 64-column tile)
    static constexpr bool STAGED_EPI = (NS == 2);
    static constexpr int EPI_BYTES = STAGED_EPI ? EPI_BUFS * EPI_WARPS * 32 * 32 * 4 : 0;        // 32 / 64 KB
    static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + EPI_BYTES + 1024 /*align*/ + 256 /*barriers*/;
    // kind::f16 instruction descriptor: D=f32 (bit 4), A/B format (0 = f16, 1 = bf16) at bits 7/10, K-major, N=256, M=128
    static constexpr uint32_t FMT = (NS == 2) ? 0u : 1u;
    static constexpr uint32_t IDESC = (1u << 4) | (FMT << 7) | (FMT << 10) | ((uint32_t)(BNT >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
};
static_assert(Cfg<2>::SMEM_BYTES <= 232448 && Cfg<2, 1>::SMEM_BYTES <= 232448 && Cfg<3>::SMEM_BYTES <= 232448 && Cfg<2, 1, 64>::SMEM_BYTES <= 232448,
              "exceeds 227 KB of shared memory");

struct Params {
    int M, N, K;                 // N % 256 == 0, K % 32 == 0
    int n_switch;                // n-tiles >= n_switch read mapA1 instead of mapA0
    int n_off;                   // column window: this launch computes output columns [n_off, n_off + N) of the full projection
                                 // (W rows, bias entries and output columns are all shifted by n_off; multiple of 32)
    float out_scale;             // accumulator scale (undoes the power-of-two weight pre-scaling of the fp16 format)
    const float* bias;           // [N] or null
    float* C; int ldc;           // fp32 output (or null)
    const float* R; int ldr;     // residual added to the fp32 output (may alias C) or null
    uint16_t* Cs; long long cs_split_stride; int ldcs;   // split output [NS][*][ldcs] in the operand format (or null)
    int relu;
    int* overflow;               // set to 1 if a split output exceeds the fp16 range (NS = 2)
    unsigned stagger_ns;         // start delay per phase group (blockIdx & 3), 0 = none
    int tma_out;                 // NS = 2: write C through mapC with TMA bulk stores (reduce-add when R aliases C)
    int dry_store;               // benchmark only: format and stage the output but do not issue the TMA stores
    const int* stop;
};

// ---- PTX wrappers ---------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t}"
        ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2) {
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                 ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* map, uint32_t src, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
                 ::"l"(map), "r"(src), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma_reduce_add_2d(const CUtensorMap* map, uint32_t src, int c0, int c1) {
    asm volatile("cp.reduce.async.bulk.tensor.2d.global.shared::cta.add.bulk_group [%0, {%2, %3}], [%1];"
                 ::"l"(map), "r"(src), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma_store_3d(const CUtensorMap* map, uint32_t src, int c0, int c1, int c2) {
    asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];"
                 ::"l"(map), "r"(src), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N = 0> __device__ __forceinline__ void tma_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void tma_store_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tmem_alloc(uint32_t dst_smem, uint32_t cols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t cols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(cols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void umma_bf16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
          "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
          "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
          "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// K-major operand tile, 64-byte rows, SWIZZLE_64B: 8-row groups are 512 B apart (SBO); LBO unused.
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);       // start address      bits [ 0,14)
    d |= (uint64_t)1 << 16;                         // leading byte offset bits [16,30) (ignored for swizzled K-major)
    d |= (uint64_t)(512 >> 4) << 32;                // stride byte offset  bits [32,46)
    d |= (uint64_t)1 << 46;                         // descriptor version  bits [46,48) = 1 on sm_100
    d |= (uint64_t)4 << 61;                         // layout type         bits [61,64): 4 = SWIZZLE_64B
    return d;
}

// ---- the kernel -----------------------------------------------------------------------------------
template <int NS, int V = 0, int BNT = BN>
__global__ void __launch_bounds__(NUM_THREADS, 1)
gemm_kernel(const __grid_constant__ CUtensorMap mapA0, const __grid_constant__ CUtensorMap mapA1,
            const __grid_constant__ CUtensorMap mapW, const __grid_constant__ CUtensorMap mapC, const Params p) {
    using C_ = Cfg<NS, V, BNT>;
    constexpr int STAGES = C_::STAGES, DRAIN_KB = C_::DRAIN_KB, STAGE_BYTES = C_::STAGE_BYTES;
    constexpr int HALF = BNT / 2;                 // columns per epilogue warp (128 or 32)
    constexpr int NJ = HALF / 32;                 // 32-column blocks per epilogue warp (4 or 1)
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");      // programmatic dependent launch (see FFB_PDL_SYNC): the next kernel's CTAs
                                                                         // may be scheduled as ours retire
    // everything up to the first read of global memory (barrier init, TMEM allocation) overlaps the previous kernel's tail
    bool stopped = false;

    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t epi_base = smem_base + STAGES * STAGE_BYTES;
    const uint32_t bar_base = epi_base + C_::EPI_BYTES;
    // barriers: full[STAGES], empty[STAGES], tfull[2], tempty[2]; then the TMEM base address slot
    auto full_bar = [&](int s) { return bar_base + 8u * s; };
    auto empty_bar = [&](int s) { return bar_base + 8u * (STAGES + s); };
    auto tfull_bar = [&](int b) { return bar_base + 8u * (2 * STAGES + b); };
    auto tempty_bar = [&](int b) { return bar_base + 8u * (2 * STAGES + 2 + b); };
    const uint32_t tmem_slot = bar_base + 8u * (2 * STAGES + 4);
    uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));          // generic pointer to the aligned base
    uint32_t* tmem_slot_ptr = reinterpret_cast<uint32_t*>(smem_gen + (tmem_slot - smem_base));

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int m_tiles = (p.M + BM - 1) / BM, n_tiles = p.N / BNT, k_chunks = p.K / BK;
    const int num_tiles = m_tiles * n_tiles;

    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
        for (int b = 0; b < 2; ++b) { mbar_init(tfull_bar(b), 1); mbar_init(tempty_bar(b), EPI_WARPS); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 2) tmem_alloc(tmem_slot, TMEM_COLS);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot_ptr;
    asm volatile("griddepcontrol.wait;" ::: "memory");                    // the previous kernel has completed and its writes are visible
    stopped = (p.stop != nullptr && *p.stop != 0);                         // early-stopped decode: skip the work, still release TMEM below

    // De-phase the persistent CTAs: with identical tiles they would all reach their store phase together and hit the
    // HBM write path as one burst while the tensor pipe idles, then all compute while the write path idles
    // (measured: store time ADDED to mainloop time).  Four phase groups spread the bursts over a tile period.
    if (p.stagger_ns > 0) __nanosleep((blockIdx.x & 3u) * p.stagger_ns);

    if (stopped) {
    } else if (warp < EPI_WARP0) {
        asm volatile("setmaxnreg.dec.sync.aligned.u32 40;");
        if (warp == 0 && lane == 0) {
            // ===== TMA producer =====
            int stage = 0; uint32_t phase = 0;
            for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
                const int mt = t / n_tiles, nt = t % n_tiles;
                const CUtensorMap* mapA = (nt >= p.n_switch) ? &mapA1 : &mapA0;
                for (int kc = 0; kc < k_chunks; ++kc) {
                    mbar_wait(empty_bar(stage), phase ^ 1u);
                    mbar_expect_tx(full_bar(stage), STAGE_BYTES);
                    const uint32_t sa = smem_base + stage * STAGE_BYTES;
                    const uint32_t sb = sa + NS * A_TILE_BYTES;
#pragma unroll
                    for (int s = 0; s < NS; ++s) tma_load_3d(sa + s * A_TILE_BYTES, mapA, full_bar(stage), kc * BK, mt * BM, s);
#pragma unroll
                    for (int s = 0; s < NS; ++s) tma_load_3d(sb + s * C_::B_TILE, &mapW, full_bar(stage), kc * BK, nt * BNT + p.n_off, s);
                    if (++stage == STAGES) { stage = 0; phase ^= 1u; }
                }
            }
        } else if (warp == 1 && lane == 0) {
            // ===== MMA issuer =====
            int stage = 0; uint32_t phase = 0; uint32_t c = 0;            // c counts drains (accumulator uses)
            for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
                for (int kc = 0; kc < k_chunks; ++kc) {
                    const uint32_t buf = c & 1u;
                    const bool first_in_drain = (kc % DRAIN_KB) == 0;
                    const bool last_in_drain = ((kc % DRAIN_KB) == DRAIN_KB - 1) || (kc == k_chunks - 1);
                    if (first_in_drain) mbar_wait(tempty_bar(buf), ((c >> 1) & 1u) ^ 1u);   // epilogue drained this accumulator
                    mbar_wait(full_bar(stage), phase);                     // operands landed
                    tc_fence_after();
                    const uint32_t sa = smem_base + stage * STAGE_BYTES;
                    const uint32_t sb = sa + NS * A_TILE_BYTES;
                    const uint32_t d = tmem_base + buf * BNT;
                    // correction products first (small), dominant A0.W0 last
                    constexpr int pa3[6] = {0, 1, 1, 0, 2, 0}, pb3[6] = {1, 0, 1, 2, 0, 0};
                    constexpr int pa2[3] = {1, 0, 0}, pb2[3] = {0, 1, 0};
                    uint32_t acc = first_in_drain ? 0u : 1u;
#pragma unroll
                    for (int q = 0; q < C_::NPROD; ++q) {
                        const int ia = (NS == 2) ? pa2[q % 3] : pa3[q], ib = (NS == 2) ? pb2[q % 3] : pb3[q];
                        const uint64_t da = make_smem_desc(sa + ia * A_TILE_BYTES);
                        const uint64_t db = make_smem_desc(sb + ib * C_::B_TILE);
#pragma unroll
                        for (int k = 0; k < BK / 16; ++k) {               // +32 B per K=16 step inside the 64-byte swizzle row
                            umma_bf16(d, da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), C_::IDESC, acc);
                            acc = 1;
                        }
                    }
                    umma_commit(empty_bar(stage));                         // smem stage free once these MMAs retire
                    if (last_in_drain) { umma_commit(tfull_bar(buf)); ++c; }   // accumulator ready for the epilogue
                    if (++stage == STAGES) { stage = 0; phase ^= 1u; }
                }
            }
        }
    } else {
        // ===== epilogue warps =====
        asm volatile("setmaxnreg.inc.sync.aligned.u32 232;");
        const int e = warp - EPI_WARP0;
        const int q = warp & 3;                       // TMEM lane quarter this warp may access
        const int hcol = e >> 2;                      // which half of the tile's columns
        float* stg0 = reinterpret_cast<float*>(smem_gen + (epi_base - smem_base)) + e * 1024 * C_::EPI_BUFS;   // 32x32 floats per warp and buffer (NS = 2)
        uint32_t n_blk = 0;                           // 32x32 blocks handed to TMA so far (selects the staging buffer)
        float acc[HALF];
        uint32_t c = 0;
        for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
            const int mt = t / n_tiles, nt = t % n_tiles;
            const int n_drains = (k_chunks + DRAIN_KB - 1) / DRAIN_KB;
            for (int dr = 0; dr < n_drains; ++dr, ++c) {
                const uint32_t buf = c & 1u;
                mbar_wait(tfull_bar(buf), (c >> 1) & 1u);
                tc_fence_after();
                const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + buf * BNT + hcol * HALF;
                if constexpr (NJ == 4) {
#pragma unroll
                    for (int j = 0; j < 2; ++j) {                  // two 32-column loads in flight per wait
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
                } else {
                    uint32_t v0[32];
                    tmem_ld32(taddr, v0);
                    tmem_ld_wait();
                    if (dr == 0) {
#pragma unroll
                        for (int i = 0; i < 32; ++i) acc[i] = __uint_as_float(v0[i]);
                    } else {
#pragma unroll
                        for (int i = 0; i < 32; ++i) acc[i] += __uint_as_float(v0[i]);
                    }
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(tempty_bar(buf));
            }
            const int col0 = nt * BNT + hcol * HALF + p.n_off;
            const int row0 = mt * BM + q * 32;
            if constexpr (C_::STAGED_EPI) {
                // ---- coalesced epilogue: 32x32 blocks through a swizzled per-warp smem buffer; in the read phase a lane owns one column ----
                const bool tma_out = p.tma_out && (p.C != nullptr);
                const bool tma_cs = p.tma_out && (p.C == nullptr) && (p.Cs != nullptr);
#pragma unroll
                for (int j = 0; j < NJ; ++j) {                // (fully unrolled: acc[] must stay in registers)
                    const uint32_t sbuf = (C_::EPI_BUFS == 2) ? (n_blk & 1u) : 0u;
                    float* stg = stg0 + sbuf * 1024;
                    const uint32_t stg_s = epi_base + ((uint32_t)e * C_::EPI_BUFS + sbuf) * 4096u;
                    if (tma_cs) {
                        // ---- asynchronous split output: the block goes out as two fp16 tiles (hi, lo) of 32 x 32 halves, laid out as the
                        // SWIZZLE_64B boxes of mapC (3-D: col, row, split) expect: 64-byte rows, 16-byte chunk c stored at c ^ ((row >> 1) & 3)
                        if (lane == 0) tma_store_wait_read<C_::EPI_BUFS - 1>();
                        __syncwarp();
                        ++n_blk;
                        float amax = 0.f;                                // max |value| of the block: one FMNMX per value instead of two compares
#pragma unroll
                        for (int cc = 0; cc < 4; ++cc) {                 // 8 values per 16-byte chunk
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
                        const bool ovf = !(amax <= 65504.f);              // +-inf included (a NaN would already be a NaN in the reference)
                        if (ovf && p.overflow) *p.overflow = 1;
                        fence_proxy_async_smem();
                        __syncwarp();
                        if (lane == 0 && !p.dry_store) {
                            tma_store_3d(&mapC, stg_s, col0 + j * 32, row0, 0);
                            tma_store_3d(&mapC, stg_s + 2048u, col0 + j * 32, row0, 1);
                            tma_store_commit();
                        }
                        continue;
                    }
                    if (tma_out) {
                        // ---- asynchronous path: the 32x32 block (value = acc*scale + bias, ReLU) is laid out exactly as the
                        // SWIZZLE_128B box of mapC expects and handed to the TMA engine: plain store, or reduce-add into C when the
                        // residual aliases C (x += ...).  The warp does not wait for global memory, only for its staging buffer.
                        if (lane == 0) tma_store_wait_read<C_::EPI_BUFS - 1>();   // the block that used this staging buffer has been read out
                        __syncwarp();
                        ++n_blk;
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
                        fence_proxy_async_smem();                      // generic-proxy smem writes -> visible to the async proxy
                        __syncwarp();
                        if (lane == 0 && !p.dry_store) {
                            if (p.R) tma_reduce_add_2d(&mapC, stg_s, col0 + j * 32, row0);
                            else tma_store_2d(&mapC, stg_s, col0 + j * 32, row0);
                            tma_store_commit();
                        }
                        continue;
                    }
#pragma unroll
                    for (int cc = 0; cc < 8; ++cc) {
                        const float4 v = make_float4(acc[j * 32 + 4 * cc] * p.out_scale, acc[j * 32 + 4 * cc + 1] * p.out_scale,
                                                     acc[j * 32 + 4 * cc + 2] * p.out_scale, acc[j * 32 + 4 * cc + 3] * p.out_scale);
                        *reinterpret_cast<float4*>(stg + lane * 32 + ((cc ^ (lane & 7)) << 2)) = v;
                    }
                    __syncwarp();
                    const int col = col0 + j * 32 + lane;
                    const float bv = p.bias ? __ldg(p.bias + col) : 0.f;
                    const int nrows = min(32, p.M - row0);
                    if (p.C != nullptr) {
                        // R may alias C (in-place residual add): every lane reads exactly the addresses it writes below, so all
                        // 32 residual loads are issued up front (the compiler cannot hoist them past the stores by itself)
                        float rv[32];
                        if (p.R) {
#pragma unroll
                            for (int r = 0; r < 32; ++r) rv[r] = (r < nrows) ? p.R[(size_t)(row0 + r) * p.ldr + col] : 0.f;
                        }
#pragma unroll
                        for (int r = 0; r < 32; ++r) {
                            if (r < nrows) {
                                float v = stg[r * 32 + ((((lane >> 2) ^ (r & 7))) << 2) + (lane & 3)] + bv;
                                if (p.relu) v = fmaxf(v, 0.f);
                                if (p.R) v = rv[r] + v;
                                p.C[(size_t)(row0 + r) * p.ldc + col] = v;
                            }
                        }
                    }
                    if (p.Cs != nullptr) {
                        __half* s0 = reinterpret_cast<__half*>(p.Cs);
                        bool ovf = false;
#pragma unroll 8
                        for (int r = 0; r < 32; ++r) {
                            if (r < nrows) {
                                float v = stg[r * 32 + ((((lane >> 2) ^ (r & 7))) << 2) + (lane & 3)] + bv;
                                if (p.relu) v = fmaxf(v, 0.f);
                                ovf |= !(fabsf(v) <= 65504.f);
                                const __half h = __float2half_rn(v);
                                const __half l = __float2half_rn(v - __half2float(h));
                                const size_t off = (size_t)(row0 + r) * p.ldcs + col;
                                s0[off] = h;
                                s0[off + p.cs_split_stride] = l;
                            }
                        }
                        if (ovf && p.overflow) *p.overflow = 1;
                    }
                    __syncwarp();
                }
            } else {
                // ---- direct epilogue: one row per thread (un-coalesced) ----
                const int row = row0 + lane;
                if (row < p.M) {
                    if (p.C != nullptr) {
                        float* crow = p.C + (size_t)row * p.ldc + col0;
                        const float* rrow = p.R ? p.R + (size_t)row * p.ldr + col0 : nullptr;
#pragma unroll
                        for (int i = 0; i < HALF; i += 4) {
                            float4 v = make_float4(acc[i] * p.out_scale, acc[i + 1] * p.out_scale, acc[i + 2] * p.out_scale, acc[i + 3] * p.out_scale);
                            if (p.bias) {
                                const float4 b = __ldg(reinterpret_cast<const float4*>(p.bias + col0 + i));
                                v.x += b.x; v.y += b.y; v.z += b.z; v.w += b.w;
                            }
                            if (p.relu) { v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f); }
                            if (rrow) {
                                const float4 r = *reinterpret_cast<const float4*>(rrow + i);
                                v.x = r.x + v.x; v.y = r.y + v.y; v.z = r.z + v.z; v.w = r.w + v.w;
                            }
                            *reinterpret_cast<float4*>(crow + i) = v;
                        }
                    }
                    if (p.Cs != nullptr) {
                        __nv_bfloat16* s0 = reinterpret_cast<__nv_bfloat16*>(p.Cs) + (size_t)row * p.ldcs + col0;
#pragma unroll
                        for (int i = 0; i < HALF; i += 8) {
                            __align__(16) __nv_bfloat16 o0[8], o1[8], o2[8];
#pragma unroll
                            for (int u = 0; u < 8; ++u) {
                                float x = acc[i + u] * p.out_scale;
                                if (p.bias) x += __ldg(p.bias + col0 + i + u);
                                if (p.relu) x = fmaxf(x, 0.f);
                                split3_bf16(x, o0[u], o1[u], o2[u]);
                            }
                            *reinterpret_cast<uint4*>(s0 + i) = *reinterpret_cast<const uint4*>(o0);
                            *reinterpret_cast<uint4*>(s0 + p.cs_split_stride + i) = *reinterpret_cast<const uint4*>(o1);
                            *reinterpret_cast<uint4*>(s0 + 2 * p.cs_split_stride + i) = *reinterpret_cast<const uint4*>(o2);
                        }
                    }
                }
            }
        }
    }
    if (NS == 2 && warp >= EPI_WARP0 && lane == 0) tma_store_wait_all();   // bulk stores must finish before the CTA exits
    tc_fence_before();
    __syncthreads();
    if (warp == 2) tmem_dealloc(tmem_base, TMEM_COLS);
}

}  // namespace tc
}  // namespace ffb
