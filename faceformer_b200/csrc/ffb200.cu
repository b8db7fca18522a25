// ffb200.cu -- host side of libffb200.so: handle, weight layout, batch planning, launch sequences
// and the extern "C" entry points declared in include/ffb200.h.
//
// Data layout in HBM (all fp32 unless noted):
//   weights   one blob in reference state_dict order (+ per-decoder-layer cross K/V weights re-packed
//             as [Ld*E, E] so that all layers' cross-attention K (and V) come from ONE GEMM each)
//   mem       packed edge memory [R, E], R = sum_i (n_valid_i + 4): only un-masked rows exist, so the
//             key-padding mask of the reference (transformer.py:170,250) is realised structurally
//   Kc, Vc    cross-attention K/V cache [R, Ld*E]: computed ONCE per wireframe (the reference
//             re-projects them every step for every sequence, transformer.py:248-251)
//   tok       int32 [T, B_eff] step-major token buffer (== `predicts`, model_para.py:207,229)
//   x, x2, qkv, att, h   fp32 activations of the current step, rows ordered (sequence, position)
//   a_x2, a_x2p, a_att, a_h, a_qkv, a_qc, kc_h, vc_h   the same tensors as 16-bit operand splits (fp16x2: [2][rows][cols]) for the
//             tensor-core pipeline: written by LayerNorm / GEMM / attention epilogues, read by TMA
#include "../../include/ffb200.h"
#include "kernels.cuh"
#include "gemm_tc.cuh"
#include "gemm_tc2.cuh"
#include "attn_mma.cuh"
#include "attn_f16.cuh"
#include "attn_h.cuh"
#include "attn_x.cuh"
#include "enc64.cuh"
#include "attn_l.cuh"
#include "persist.cuh"

#include <algorithm>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <new>
#include <string>
#include <unordered_map>
#include <vector>

using namespace ffb;

namespace {

thread_local std::string g_create_error;

struct DevBuf {
    void* p = nullptr;
    size_t cap = 0;
    cudaError_t ensure(size_t bytes) {
        if (bytes <= cap) return cudaSuccess;
        if (p) { cudaError_t e = cudaFree(p); p = nullptr; cap = 0; if (e != cudaSuccess) return e; }
        size_t want = bytes + bytes / 8 + 256;
        cudaError_t e = cudaMalloc(&p, want);
        if (e != cudaSuccess) { p = nullptr; cap = 0; return e; }
        cap = want;
        // fresh memory is zeroed: TMA boxes and GEMM tiles may touch rows past the valid ones, and 0 x (stale NaN pattern) must not happen
        return cudaMemset(p, 0, want);
    }
    void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
    template <class T> T* as() const { return reinterpret_cast<T*>(p); }
};

struct AttnW { const float *in_w, *in_b, *out_w, *out_b; };
struct EncLayerW { AttnW sa; const float *l1w, *l1b, *l2w, *l2b, *n1w, *n1b, *n2w, *n2b; };
struct DecLayerW { AttnW sa, ca; const float *l1w, *l1b, *l2w, *l2b, *n1w, *n1b, *n2w, *n2b, *n3w, *n3b; };

struct Weights {
    const float *tok_table, *e0w, *e0b, *e2w, *e2b, *pos, *qpos;
    std::vector<EncLayerW> enc;
    const float *enc_nw, *enc_nb;
    std::vector<DecLayerW> dec;
    const float *dec_nw, *dec_nb, *proj_w, *proj_b;
    const float *ckw, *ckb, *cvw, *cvb;      // packed cross K / V weights of all decoder layers
};

}  // namespace

struct ffb_handle {
    ffb_config cfg;
    int E, H, FF, L, T, Le, Ld;
    std::string err;
    bool weights_loaded = false, encoded = false, decoded = false;
    int opt_dedup = 1, opt_prune = 1, opt_timing = 0;
    int opt_enc_prec = 2;                         // encoder + cross K/V projections: 2 = float64 (default, enc64.cuh), 0 = fp16x2 tcgen05 / fp32 SIMT
    int opt_head64 = 1;                           // decoder.norm + project + pointer dot of the last position in float64
    DevBuf x64, y64, yp64, qkv64, att64, h64;
    int opt_persist = 1;                          // whole greedy loop as one persistent cooperative kernel (persist.cuh): 0 off, 1 auto (small batches), 2 wherever supported
    int pd_grid = 0;                              // co-resident CTAs of decode_persistent_kernel on this device (0: cooperative launch unavailable)
    DevBuf pd_sync, pd_prof;                      // unsigned barrier counter + int[T] per-step counters; FFB_PD_PROF clock sums
    bool used_persist = false;                    // the last decode ran (at least its first steps) in the persistent kernel
    bool l0_suspend = false;                      // ... and the per-step kernels behind it must not read the layer-0 cache
    const unsigned char* train_kmask = nullptr;   // non-null while ffb_forward_train runs its decoder pass: label padding mask [B * (T - 1)]
    DevBuf d_label, d_label_mask, d_kmask, dn_q_begin, dn_edge_dst, dn_pos_idx;
    int opt_l0cache = 1;                          // decoder layer 0: q / k / v of earlier prefix positions are cached (exact), only the new position is projected
    DevBuf qkv0_cache, a_qkv0; CUtensorMap ms_qkv0; bool l0_ok = false;
    DevBuf e0pad, a_c; CUtensorMap ms_x2;         // W0 zero-padded to [E, 128]; coordinates as fp16x2 operand [2][cap][128]; split-store map of a_x2
    DevBuf projT, memW, hy32;                     // folded head: [W_project^T ; b_project] (E+1 x E), memory . projT^T [R, E+4], LN of the last position [B, E]
    bool enc_used_64 = false;
    DevBuf d_tile_off;                            // [N + 1] prefix of ceil(vlen / 256): work items (pairs of query tiles) of attn_l.cuh
    int opt_attn_long = 1;                        // encoder self-attention with > 256 keys per wireframe on the tcgen05 streaming kernel (0: mma.sync)
    int opt_encode_only = 0;                      // encoder-only use (BASELINE configs[4]): no decode workspaces, no cross K/V cache, no folded head
    bool encode_only = false;                     // ... of the encoded batch
    int opt_force_F = 0;                          // F of the GLOBAL batch when this handle decodes a share of it (0 = local max(num_input))
    // cross-rank stop predicate (batch splitting): flag buffers of all ranks, mapped through CUDA IPC
    bool xchg_on = false; int xchg_rank = 0, xchg_world = 1, xchg_epoch = 0;
    int* xchg_buf = nullptr; int* xchg_peers[XCHG_MAXW] = {};
    int opt_beam = 1;                             // beam width W (parallel mode; 1 = greedy): every anchor owns W consecutive sequences
    int W = 1;                                    // beam width of the encoded batch
    int tok_sel = 0;                              // which half of the double-buffered token array is current (beam re-ordering)
    DevBuf beam_cum;                              // double [B]: cumulative log-probabilities of the hypotheses
    int64_t launches = 0;

    DevBuf wblob, wcross;
    Weights w;

    // batch plan (host)
    int N = 0, F = 0;
    long long B_full = 0, B = 0, R = 0, Re = 0;
    int max_vlen = 0, max_seq_per_wf = 0;
    std::vector<int> h_row_off, h_vlen, h_seq_off;
    // batch plan (device)
    DevBuf d_row_off, d_vlen, d_pos_idx, d_edge_src, d_edge_dst, d_seq_wf, d_seq_first, d_seq_off, d_slot_seq, d_seq_slot;
    // staging for host-located inputs/outputs
    DevBuf d_coords, d_predict, d_out_stage, d_mask_stage, d_prefix;
    // persistent per batch
    DevBuf mem, Kc, Vc, tok, logits, state;     // state: int[4] = stop, steps_run, nonstop_count, eos_count
    // activations
    DevBuf x, x2, qkv, att, hb, xl;
    int last_P = 0;                               // P of the last executed pointer projection (seq2seq 'pointer')
    // tensor-core path (gemm_tc.cuh): fp16x2 / bf16x3 split weights + activation operands, TMA tensor maps
    bool tc_ok = false;                           // geometry supported (E, FF multiples of 256)
    int opt_tc = 1;                               // 0 off, 1 auto (M >= TC_MIN_ROWS), 2 force
    int opt_stagger = 0;                          // de-phase persistent GEMM CTAs (measured: no effect; kept for experiments)
    int opt_tma_out = 1;                          // fp16x2 GEMM: asynchronous TMA store / reduce-add epilogue
    int opt_gemm_variant = 2;                     // fp16x2 GEMM: 0 / 1 = single-CTA pipeline variant (gemm_tc.cuh Cfg<NS, V>), 2 = variant per launch,
                                                  // 3 = CTA-pair kernel (gemm_tc2.cuh, cta_group::2) wherever the TMA epilogue applies
    std::unordered_map<const CUtensorMap*, CUtensorMap> w_half_maps;   // W map (256-row boxes) -> its twin with 128-row boxes (CTA-pair kernel)
    std::unordered_map<const CUtensorMap*, CUtensorMap> w_skinny_maps; // ... and with 64-row boxes (64-column tiles for small M)
    int opt_skinny = 1;                           // fp16x2 GEMMs whose 128 x 256 tiles would leave most SMs idle run with 128 x 64 tiles
    CUtensorMap mc_x, mc_xl, mc_qkv3, mc_qkv1, mc_att;   // fp32 output maps of the decode-step activation buffers
    CUtensorMap ms_h;                             // fp16x2 split STORE map of a_h (FFN hidden)
    // "half pipeline" (fp16x2 GEMM + fp16x2 attention): q,k,v and the cross-attention query never exist in fp32
    DevBuf a_qkv, a_qc, kc_h, vc_h;               // [2][cap][3E], [2][cap][E], [2][R][Ld*E] halves
    DevBuf a_ql; CUtensorMap ms_ql; long long cap_b = 0;   // [2][cap_b][E]: q of the last prefix position (pruned last layer)
    CUtensorMap ms_qkv, ms_qc;
    bool half_pipe = false;                       // set per batch at plan time
    // tcgen05 attention (attn_x.cuh): TMA load maps
    CUtensorMap mx_q, mx_k, mx_vrow;              // cross: Q rows of a_qc, K rows of kc_h, V rows of vc_h (MN-major operand)
    CUtensorMap msf_q, msf_k, msf_v;              // self: q / k / v sections of a_qkv
    bool attn_x_ok = false;                       // set per batch: half pipeline, <= 256 keys per wireframe, <= 255 wireframes
    int opt_attn_x = 3;                           // bit 0: cross-attention on tcgen05, bit 1: self-attention on tcgen05
    int opt_attn_mma = 2;                         // attention core: 2 = mma.sync fp16x2 kernel (decode, while the GEMM format is fp16x2),
                                                  // 1 = mma.sync 3xTF32 kernel, 0 = fp32 SIMT kernels
    bool attn_allow_f16 = true;                   // cleared while encoding (an overflow flag raised there would be lost)
    int ovf_slot = 4;                             // state[] slot the fp16-range checks of the kernels raise: 4 = decode, 6 = tensor-core encoder
    int opt_pointer_batched = 1;                  // pointer head batched per wireframe (bit-identical to pointer_kernel)
    int opt_pdl = 2;                              // programmatic dependent launch for the decode-step kernels: 0 off, 1 on, 2 auto (small batches:
                                                  // measured +5..10 % at N = 1, -3 % on the 32-wireframe bench batch)
    bool pdl_on = false;                          // decided per batch
    int opt_enc_tc = 1;                           // encoder layers + cross K/V projections on the tcgen05 pipeline when the batch allows it
    bool enc_used_tc = false;
    CUtensorMap mc_kc, mc_vc;                     // fp32 output maps of the cross-attention cache
    int num_sms = 148;
    int tc_fmt = 2;                               // operand format: 2 = fp16x2 (3 MMA passes), 3 = bf16x3 (6 passes)
    int fp16_fallbacks = 0;                       // decodes re-run in bf16x3 because an activation exceeded the fp16 range
    struct DecTcW { CUtensorMap sa_in, sa_out, ca_q, ca_out, l1, l2; float s_sa_in, s_sa_out, s_ca_q, s_ca_out, s_l1, s_l2;
                    const uint16_t *p_sa_in, *p_sa_out, *p_ca_q, *p_ca_out, *p_l1, *p_l2; };     // the split arrays themselves (persist.cuh)
    struct EncTcW { CUtensorMap sa_in, sa_out, l1, l2; float s_sa_in, s_sa_out, s_l1, s_l2; };
    struct TcSet {                                // everything that depends on the operand format
        bool ready = false;
        DevBuf wsplit;                            // [fmt][N][K] 16-bit splits of every decode-step weight matrix
        std::vector<DecTcW> layers;
        std::vector<EncTcW> enc;                  // encoder layers
        CUtensorMap ck, cv; float s_ck = 1.f, s_cv = 1.f;   // packed cross-attention K / V projections of all decoder layers [Ld*E, E]
        CUtensorMap proj; float s_proj = 1.f;
        CUtensorMap emb2; float s_emb2 = 1.f;     // second linear of the value embedding (embedding.py:17), E x E
        CUtensorMap emb0; float s_emb0 = 1.f; bool has_emb0 = false;   // first linear (embedding.py:15), K = in_dim zero-padded to 128
        CUtensorMap m_c;                          // activation-operand map of a_c (padded coordinates)
        CUtensorMap m_x2, m_x2p, m_att, m_h;      // activation-operand maps (re-encoded per batch)
    };
    TcSet tcs[2];                                 // [fmt - 2]
    DevBuf a_x2, a_x2p, a_att, a_h;               // [fmt][cap_rows][E or FF] 16-bit (sized for 3 splits)
    long long cap_rows = 0;
    // per-kernel-class profiling (FFB_OPT_PROFILE): event pairs around every launch
    int opt_profile = 0;
    std::vector<cudaEvent_t> prof_pool;
    struct ProfRec { int cls; double flops; };
    std::vector<ProfRec> prof_recs;
    double sum_seq_vlen = 0, sum_vlen2 = 0;       // sum_i seqs_i*vlen_i and sum_i vlen_i^2 (attention FLOP accounting)
    cudaEvent_t ev[3] = {nullptr, nullptr, nullptr};
    volatile int* h_stop = nullptr;               // pinned host mirror of the stop flag, one slot per step (early stop without a host sync)
    int h_stop_cap = 0;
    int steps_launched = 0;                       // decode steps whose kernels were launched by the last ffb_decode_greedy
    std::vector<CUtensorMap> m_last;              // per prefix length P: strided view of a_x2p (rows b*P + P-1), pruned last layer
    std::vector<uint8_t> h_mask;
    std::vector<int64_t> h_num_input;
};

namespace {

int fail(ffb_handle* h, int code, const char* fmt, ...) {
    char buf[1024];
    va_list ap; va_start(ap, fmt); vsnprintf(buf, sizeof buf, fmt, ap); va_end(ap);
    if (h) h->err = buf; else g_create_error = buf;
    return code;
}

#define CU(h, expr) do { cudaError_t _e = (expr); if (_e != cudaSuccess) \
    return fail((h), FFB_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); } while (0)
#define FFB_TRY(expr) do { int _r = (expr); if (_r != FFB_OK) return _r; } while (0)

inline int* tok_cur(ffb_handle* h) { return h->tok.as<int>() + (size_t)h->tok_sel * h->T * h->B; }
inline int* ovf_ptr(ffb_handle* h) { return h->state.as<int>() ? h->state.as<int>() + h->ovf_slot : nullptr; }

enum { PC_LINEAR = 0, PC_LAYERNORM, PC_ATTN_ROWS, PC_ATTN_TILED, PC_POINTER, PC_OTHER, PC_LINEAR_TC, PC_COUNT };
constexpr int TC_MIN_ROWS = 2048;       // encoder: below this many memory rows the tensor-core encoder is not worth its operand formatting
constexpr int TC_MIN_ROWS_DECODE = 1;    // decode steps: the tcgen05 GEMM beats the SIMT kernel at every M (measured r2: 4x at M <= 258, profiles/probe_small_r2a.json)

inline void prof_begin(ffb_handle* h, int cls, double flops, cudaStream_t s) {
    if (!h->opt_profile) return;
    const size_t i = h->prof_recs.size();
    while (h->prof_pool.size() < 2 * (i + 1)) { cudaEvent_t e; cudaEventCreate(&e); h->prof_pool.push_back(e); }
    h->prof_recs.push_back({cls, flops});
    cudaEventRecord(h->prof_pool[2 * i], s);
}
inline void prof_end(ffb_handle* h, cudaStream_t s) {
    if (!h->opt_profile) return;
    cudaEventRecord(h->prof_pool[2 * (h->prof_recs.size() - 1) + 1], s);
}

const char* validate_config(const ffb_config* c) {
    if (!c) return "config is NULL";
    if (c->abi_version != FFB_ABI_VERSION) return "abi_version mismatch";
    if (c->mode != FFB_MODE_PARALLEL && c->mode != FFB_MODE_SEQ2SEQ) return "mode must be FFB_MODE_PARALLEL or FFB_MODE_SEQ2SEQ";
    if (c->num_model <= 0 || c->num_model % 128 != 0 || c->num_model > 1024) return "num_model must be a multiple of 128, <= 1024";
    if (c->num_head <= 0 || c->num_model != c->num_head * 64) return "head dim (num_model / num_head) must be 64";
    if (c->num_feedforward <= 0 || c->num_feedforward % 4 != 0) return "num_feedforward must be a positive multiple of 4";
    if (c->num_encoder_layers < 1 || c->num_decoder_layers < 1) return "need >= 1 encoder and decoder layer";
    if (c->in_dim <= 0 || c->in_dim % 4 != 0) return "in_dim must be a positive multiple of 4";
    if (c->num_lines < 1) return "num_lines must be >= 1";
    if (c->num_token != 4) return "num_token must be 4 (PAD,SOS,SEP,EOS)";
    if (c->seq_len < 2) return "seq_len must be >= 2";
    if (c->device < 0) return "device must be >= 0";
    return nullptr;
}

size_t weight_count(const ffb_config* c) {
    const size_t E = c->num_model, FF = c->num_feedforward, L = c->num_lines + c->num_token, T = c->seq_len;
    const size_t attn = 3 * E * E + 3 * E + E * E + E;
    const size_t ffn = FF * E + FF + E * FF + E;
    size_t n = c->num_token * E + E * c->in_dim + E + E * E + E + L * E + T * E;
    n += (size_t)c->num_encoder_layers * (attn + ffn + 4 * E) + 2 * E;
    n += (size_t)c->num_decoder_layers * (2 * attn + ffn + 6 * E) + 2 * E;
    n += E * E + E;
    return n;
}

void bind_weights(ffb_handle* h) {
    const float* p = h->wblob.as<float>();
    const size_t E = h->E, FF = h->FF, L = h->L, T = h->T, in = h->cfg.in_dim;
    auto take = [&](size_t n) { const float* r = p; p += n; return r; };
    Weights& w = h->w;
    w.tok_table = take(h->cfg.num_token * E);
    w.e0w = take(E * in); w.e0b = take(E); w.e2w = take(E * E); w.e2b = take(E);
    w.pos = take(L * E); w.qpos = take(T * E);
    auto attn = [&]() { AttnW a; a.in_w = take(3 * E * E); a.in_b = take(3 * E); a.out_w = take(E * E); a.out_b = take(E); return a; };
    w.enc.resize(h->Le);
    for (auto& l : w.enc) {
        l.sa = attn();
        l.l1w = take(FF * E); l.l1b = take(FF); l.l2w = take(E * FF); l.l2b = take(E);
        l.n1w = take(E); l.n1b = take(E); l.n2w = take(E); l.n2b = take(E);
    }
    w.enc_nw = take(E); w.enc_nb = take(E);
    w.dec.resize(h->Ld);
    for (auto& l : w.dec) {
        l.sa = attn(); l.ca = attn();
        l.l1w = take(FF * E); l.l1b = take(FF); l.l2w = take(E * FF); l.l2b = take(E);
        l.n1w = take(E); l.n1b = take(E); l.n2w = take(E); l.n2b = take(E); l.n3w = take(E); l.n3b = take(E);
    }
    w.dec_nw = take(E); w.dec_nb = take(E);
    w.proj_w = take(E * E); w.proj_b = take(E);
}

// Launch with the programmatic-stream-serialization attribute (FFB_OPT_PDL): the kernel may be scheduled before its predecessor in
// the stream has finished; every kernel launched this way starts with FFB_PDL_SYNC() / griddepcontrol.wait (kernels.cuh).
template <class... KArgs, class... Args>
inline void launch_k(ffb_handle* h, void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s, Args&&... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = h->pdl_on ? 1 : 0;
    cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

inline int grid1d(long long total, int block = 256) {
    long long g = (total + block - 1) / block;
    return (int)std::max<long long>(1, std::min<long long>(g, 148LL * 32));
}

// ---- launch helpers -------------------------------------------------------------------------------
struct Lin {
    const float* A = nullptr; int lda = 0; const int* a_rows = nullptr;
    const float* W = nullptr; int ldw = 0; const float* bias = nullptr;
    float* C = nullptr; int ldc = 0; const int* c_rows = nullptr;
    const float* R = nullptr; int ldr = 0;
    const float* pos = nullptr; int ldpos = 0; const int* pos_idx = nullptr; int pos_mod = 0; int pos_cols = 0;
    int M = 0, N = 0, K = 0, relu = 0;
};

int launch_linear(ffb_handle* h, const Lin& l, const int* stop, cudaStream_t s) {
    if (l.M <= 0) return FFB_OK;
    LinearArgs a;
    a.A = l.A; a.lda = l.lda; a.a_rows = l.a_rows; a.W = l.W; a.ldw = l.ldw; a.bias = l.bias;
    a.C = l.C; a.ldc = l.ldc; a.c_rows = l.c_rows; a.R = l.R; a.ldr = l.ldr;
    a.pos = l.pos; a.ldpos = l.ldpos; a.pos_idx = l.pos_idx; a.pos_mod = l.pos_mod; a.pos_cols = l.pos_cols;
    a.M = l.M; a.N = l.N; a.K = l.K; a.relu = l.relu; a.stop = stop;
    if (l.pos && l.pos_cols < l.N && (l.pos_cols % LBN) != 0)
        return fail(h, FFB_ERR_ARG, "linear: pos_cols (%d) must be a multiple of %d", l.pos_cols, LBN);
    if (l.K % 4 != 0 || l.N % 4 != 0) return fail(h, FFB_ERR_ARG, "linear: K and N must be multiples of 4");
    dim3 grid((l.M + LBM - 1) / LBM, (l.N + LBN - 1) / LBN);
    prof_begin(h, PC_LINEAR, 2.0 * l.M * (double)l.N * l.K, s);
    linear_kernel<<<grid, 256, 0, s>>>(a);
    prof_end(h, s);
    h->launches++;
    CU(h, cudaGetLastError());
    return FFB_OK;
}

int launch_ln(ffb_handle* h, const float* x, const float* g, const float* b, float* y, int M, int E, const int* stop, cudaStream_t s) {
    if (M <= 0) return FFB_OK;
    prof_begin(h, PC_LAYERNORM, 8.0 * M * (double)E, s);
    launch_k(h, layernorm_kernel, dim3((M + 7) / 8), dim3(256), 0, s, x, g, b, y, M, E, stop);
    prof_end(h, s);
    h->launches++;
    CU(h, cudaGetLastError());
    return FFB_OK;
}

int launch_attn_mma(ffb_handle* h, const float* Q, int ldq, const float* K, const float* V, int ldk, float* O, int ldo,
                    const AttnGroups& g, int G, int max_q_rows, double qk_pairs, int prof_class, const int* stop, cudaStream_t s,
                    uint16_t* Os, long long os_stride);

int launch_attn_rows(ffb_handle* h, const float* Q, int ldq, const float* K, const float* V, int ldk, float* O, int ldo,
                     int G, int nq, int nk, int q_stride, int q_off, int k_stride, int o_stride, const int* stop, cudaStream_t s,
                     uint16_t* Os = nullptr, long long os_stride = 0, const unsigned char* key_mask = nullptr, int causal = 0) {
    if (G <= 0 || nq <= 0) return FFB_OK;
    AttnGroups g{}; g.ragged = 0; g.nq = nq; g.nk = nk; g.q_stride = q_stride; g.q_off = q_off; g.k_stride = k_stride; g.o_stride = o_stride;
    g.split_fmt = h->tc_fmt; g.overflow = ovf_ptr(h); g.key_mask = key_mask; g.causal = causal;
    if (h->opt_attn_mma && !key_mask && !causal)       // (the masks of the teacher-forced pass exist in the SIMT kernel only)
        return launch_attn_mma(h, Q, ldq, K, V, ldk, O, ldo, g, G, nq, (double)G * nq * nk, PC_ATTN_ROWS, stop, s, Os, os_stride);
    dim3 grid(G, h->H, (nq + AR_BQ - 1) / AR_BQ);
    prof_begin(h, PC_ATTN_ROWS, 4.0 * 64 * (double)G * h->H * nq * nk, s);
    attn_rows_kernel<<<grid, 128, 0, s>>>(Q, ldq, K, V, ldk, O, ldo, Os, os_stride, g, stop);
    prof_end(h, s);
    h->launches++;
    CU(h, cudaGetLastError());
    return FFB_OK;
}

int launch_attn_mma(ffb_handle* h, const float* Q, int ldq, const float* K, const float* V, int ldk, float* O, int ldo,
                    const AttnGroups& g, int G, int max_q_rows, double qk_pairs, int prof_class, const int* stop, cudaStream_t s,
                    uint16_t* Os, long long os_stride) {
    if (G <= 0 || max_q_rows <= 0) return FFB_OK;
    const int qtiles = (max_q_rows + AM_BQ - 1) / AM_BQ;
    if (qtiles > 65535) return fail(h, FFB_ERR_ARG, "attention: too many query tiles per group");
    dim3 grid(G, h->H, qtiles);
    // fp16x2 attention only as part of the tensor-core pipeline; geometries that stay on the fp32 SIMT GEMMs keep the 3xTF32 kernel
    const bool f16 = h->opt_attn_mma == 2 && h->tc_fmt == 2 && h->attn_allow_f16 && h->tc_ok && h->opt_tc;
    prof_begin(h, prof_class, 4.0 * 64 * h->H * qk_pairs, s);
    if (f16) attn_f16_kernel<<<grid, 128, AF_SMEM_BYTES, s>>>(Q, ldq, K, V, ldk, O, ldo, Os, os_stride, g, stop);
    else attn_mma_kernel<<<grid, 128, AM_SMEM_BYTES, s>>>(Q, ldq, K, V, ldk, O, ldo, Os, os_stride, g, stop);
    prof_end(h, s);
    h->launches++;
    CU(h, cudaGetLastError());
    return FFB_OK;
}

int launch_attn_h(ffb_handle* h, const AttnHalfIn& in, uint16_t* Os, long long os_stride, AttnGroups g, int G, int max_q_rows, int max_nk,
                  double qk_pairs, int prof_class, const int* stop, cudaStream_t s, float* O = nullptr) {
    if (G <= 0 || max_q_rows <= 0) return FFB_OK;
    const int qtiles = (max_q_rows + AF_BQ - 1) / AF_BQ;
    if (qtiles > 65535) return fail(h, FFB_ERR_ARG, "attention: too many query tiles per group");
    g.split_fmt = h->tc_fmt; g.overflow = ovf_ptr(h);
    dim3 grid(G, h->H, qtiles);
    prof_begin(h, prof_class, 4.0 * 64 * h->H * qk_pairs, s);
    const int qr_cap = std::min(AF_BQ, (std::min(max_q_rows, AF_BQ) + 15) & ~15);      // rows per Q tile buffer
    const int kr_cap = std::min(AF_BK, (std::min(std::max(max_nk, 1), AF_BK) + 15) & ~15);   // rows per K / V tile buffer
    const int smem = (2 * qr_cap + 4 * kr_cap) * AF_S * 2;
    launch_k(h, attn_h_kernel, grid, dim3(32 * (qr_cap / 16)), (size_t)smem, s, in, O, h->E, Os, os_stride, g, qr_cap, kr_cap, stop);
    prof_end(h, s);
    h->launches++;
    CU(h, cudaGetLastError());
    return FFB_OK;
}

// tcgen05 attention (attn_x.cuh).  The caller fills the mode-specific fields of `p`; total_items / grid are derived here.
int launch_attn_x(ffb_handle* h, const CUtensorMap& mq, const CUtensorMap& mk, const CUtensorMap& mv, ax::Params p,
                  const std::vector<int>* seq_off_host, double qk_pairs, int prof_class, const int* stop, cudaStream_t s) {
    long long tiles = 0;
    if (p.mode == 0) {
        for (int i = 0; i < p.n_groups; ++i)
            tiles += ((long long)((*seq_off_host)[i + 1] - (*seq_off_host)[i]) * p.q_mul + ax::BQ - 1) / ax::BQ;
    } else {
        p.seqs_per_tile = ax::BQ / p.P;
        tiles = (p.n_seqs + p.seqs_per_tile - 1) / p.seqs_per_tile;
    }
    const long long items = tiles * h->H;
    if (items <= 0) return FFB_OK;
    if (items > 0x7fffffffLL) return fail(h, FFB_ERR_ARG, "attention: too many work items");
    p.n_heads = h->H; p.total_items = (int)items; p.stop = stop;
    const int grid = (int)std::min<long long>(items, h->num_sms);
    prof_begin(h, prof_class, 4.0 * 64 * h->H * qk_pairs, s);
    launch_k(h, ax::attn_x_kernel, dim3(grid), dim3(ax::NUM_THREADS), (size_t)ax::SMEM_BYTES, s, mq, mk, mv, p);
    prof_end(h, s);
    h->launches++;
    CU(h, cudaGetLastError());
    return FFB_OK;
}

// tcgen05 attention over many keys (attn_l.cuh): queries = keys = the rows of each group
int launch_attn_long(ffb_handle* h, const CUtensorMap& mq, const CUtensorMap& mk, const CUtensorMap& mv, al::Params p, long long total_tiles,
                     double qk_pairs, int prof_class, cudaStream_t s) {
    const long long items = total_tiles * h->H;
    if (items <= 0) return FFB_OK;
    if (items > 0x7fffffffLL) return fail(h, FFB_ERR_ARG, "attention: too many work items");
    p.n_heads = h->H; p.total_items = (int)items;
    const int grid = (int)std::min<long long>(items, h->num_sms);
    prof_begin(h, prof_class, 4.0 * 64 * h->H * qk_pairs, s);
    al::attn_long_kernel<<<grid, al::NUM_THREADS, al::SMEM_BYTES, s>>>(mq, mk, mv, p);
    prof_end(h, s);
    h->launches++;
    CU(h, cudaGetLastError());
    return FFB_OK;
}

int launch_attn_tiled(ffb_handle* h, const float* Q, int ldq, const float* K, const float* V, int ldk, float* O, int ldo,
                      const AttnGroups& g_in, int G, int max_q_rows, double qk_pairs, const int* stop, cudaStream_t s,
                      uint16_t* Os = nullptr, long long os_stride = 0) {
    if (G <= 0 || max_q_rows <= 0) return FFB_OK;
    AttnGroups g = g_in;
    g.split_fmt = h->tc_fmt; g.overflow = ovf_ptr(h);
    if (h->opt_attn_mma)
        return launch_attn_mma(h, Q, ldq, K, V, ldk, O, ldo, g, G, max_q_rows, qk_pairs, PC_ATTN_TILED, stop, s, Os, os_stride);
    if (G > 65535) return fail(h, FFB_ERR_ARG, "attention: more than 65535 groups");
    dim3 grid((max_q_rows + AT_BQ - 1) / AT_BQ, h->H, G);
    prof_begin(h, PC_ATTN_TILED, 4.0 * 64 * h->H * qk_pairs, s);
    attn_tiled_kernel<<<grid, 128, AT_SMEM_BYTES, s>>>(Q, ldq, K, V, ldk, O, ldo, Os, os_stride, g, stop);
    prof_end(h, s);
    h->launches++;
    CU(h, cudaGetLastError());
    return FFB_OK;
}

// ---- tensor-core path helpers ---------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn g_encode_tiled = nullptr;

// 16-bit [fmt][rows][K] operand array -> 3-D map (k, row, split), box {32, box_rows, 1}, 64-byte swizzle
int encode_operand_map(ffb_handle* h, CUtensorMap* m, void* base, uint64_t K, uint64_t rows, uint32_t box_rows, int fmt) {
    if (!g_encode_tiled) return fail(h, FFB_ERR_CUDA, "cuTensorMapEncodeTiled is unavailable");
    const cuuint64_t dims[3] = {K, rows, (cuuint64_t)fmt};
    const cuuint64_t strides[2] = {K * 2, rows * K * 2};
    const cuuint32_t box[3] = {(cuuint32_t)tc::BK, box_rows, 1};
    const cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = g_encode_tiled(m, fmt == 3 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, base, dims, strides, box, estr,
                                CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                                CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(h, FFB_ERR_CUDA, "cuTensorMapEncodeTiled failed with CUresult %d (K=%llu rows=%llu)", (int)r,
                                       (unsigned long long)K, (unsigned long long)rows);
    return FFB_OK;
}

// 16-bit [2][rows][ld] fp16x2 buffer -> 3-D LOAD map (col, row, split) with an explicit box and swizzle
int encode_rows_map(ffb_handle* h, CUtensorMap* m, void* base, uint64_t ld, uint64_t rows, uint32_t box_cols, uint32_t box_rows,
                    CUtensorMapSwizzle swz) {
    if (!g_encode_tiled) return fail(h, FFB_ERR_CUDA, "cuTensorMapEncodeTiled is unavailable");
    const cuuint64_t dims[3] = {ld, rows, 2};
    const cuuint64_t strides[2] = {ld * 2, rows * ld * 2};
    const cuuint32_t box[3] = {box_cols, box_rows, 1};
    const cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = g_encode_tiled(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, swz,
                                CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(h, FFB_ERR_CUDA, "cuTensorMapEncodeTiled(rows map) failed with CUresult %d", (int)r);
    return FFB_OK;
}

// fp16x2 operand view with an explicit row stride: rows `first, first + every, ...` of a [2][cap][K] buffer (the last prefix position of
// every sequence: first = P - 1, every = P)
int encode_operand_map_strided(ffb_handle* h, CUtensorMap* m, uint16_t* base, uint64_t K, uint64_t rows, uint64_t every, uint64_t first,
                               uint64_t cap_rows, uint32_t box_rows) {
    if (!g_encode_tiled) return fail(h, FFB_ERR_CUDA, "cuTensorMapEncodeTiled is unavailable");
    const cuuint64_t dims[3] = {K, rows, 2};
    const cuuint64_t strides[2] = {every * K * 2, cap_rows * K * 2};
    const cuuint32_t box[3] = {(cuuint32_t)tc::BK, box_rows, 1};
    const cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = g_encode_tiled(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, base + first * K, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(h, FFB_ERR_CUDA, "cuTensorMapEncodeTiled(strided operand) failed with CUresult %d", (int)r);
    return FFB_OK;
}

// weight operand [fmt][rows][K]: the 256-row-box map of the single-CTA kernel plus (fp16x2) its 128-row-box twin for the CTA-pair kernel
int encode_operand_map(ffb_handle* h, CUtensorMap* m, void* base, uint64_t K, uint64_t rows, uint32_t box_rows, int fmt);
int encode_weight_maps(ffb_handle* h, CUtensorMap* m, void* base, uint64_t K, uint64_t rows, int fmt) {
    int rc = encode_operand_map(h, m, base, K, rows, tc::BN, fmt);
    if (rc != FFB_OK || fmt != 2) return rc;
    CUtensorMap half, skinny;
    rc = encode_operand_map(h, &half, base, K, rows, tc2::BN / 2, fmt);
    if (rc == FFB_OK) h->w_half_maps[m] = half;
    if (rc == FFB_OK) rc = encode_operand_map(h, &skinny, base, K, rows, 64, fmt);
    if (rc == FFB_OK) h->w_skinny_maps[m] = skinny;
    return rc;
}

// fp32 [rows, ld] output buffer -> 2-D map, box 32 x 32 floats (128-byte rows), SWIZZLE_128B (the staging layout of the epilogue)
int encode_output_map(ffb_handle* h, CUtensorMap* m, void* base, uint64_t ld, uint64_t rows) {
    if (!g_encode_tiled) return fail(h, FFB_ERR_CUDA, "cuTensorMapEncodeTiled is unavailable");
    const cuuint64_t dims[2] = {ld, rows};
    const cuuint64_t strides[1] = {ld * 4};
    const cuuint32_t box[2] = {32, 32};
    const cuuint32_t estr[2] = {1, 1};
    CUresult r = g_encode_tiled(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(h, FFB_ERR_CUDA, "cuTensorMapEncodeTiled(output) failed with CUresult %d", (int)r);
    return FFB_OK;
}

// fp16x2 split buffer [2][rows][ld] -> 3-D STORE map (col, row, split), box 32 x 32 x 1 halves, SWIZZLE_64B
int encode_split_store_map(ffb_handle* h, CUtensorMap* m, void* base, uint64_t ld, uint64_t rows) {
    if (!g_encode_tiled) return fail(h, FFB_ERR_CUDA, "cuTensorMapEncodeTiled is unavailable");
    const cuuint64_t dims[3] = {ld, rows, 2};
    const cuuint64_t strides[2] = {ld * 2, rows * ld * 2};
    const cuuint32_t box[3] = {32, 32, 1};
    const cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = g_encode_tiled(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(h, FFB_ERR_CUDA, "cuTensorMapEncodeTiled(split store) failed with CUresult %d", (int)r);
    return FFB_OK;
}

struct TcLin {
    const CUtensorMap* A0 = nullptr; const CUtensorMap* A1 = nullptr; int n_switch = 1 << 30;
    const CUtensorMap* W = nullptr; float w_scale = 1.f; const float* bias = nullptr;
    float* C = nullptr; int ldc = 0; const float* R = nullptr; int ldr = 0;
    const CUtensorMap* Cmap = nullptr;          // fp32 [rows, ldc] map of C (box 32x32, SWIZZLE_128B): enables the TMA epilogue (fp16x2)
    uint16_t* Cs = nullptr; long long cs_stride = 0; int ldcs = 0;
    int M = 0, N = 0, K = 0, relu = 0, dry_store = 0, n_off = 0;
};

int launch_tc(ffb_handle* h, const TcLin& l, const int* stop, cudaStream_t s) {
    if (l.M <= 0) return FFB_OK;
    if (l.N % tc::BN != 0 || l.K % tc::BK != 0) return fail(h, FFB_ERR_ARG, "tc gemm: N %% 256 and K %% 32 must be 0 (N=%d K=%d)", l.N, l.K);
    tc::Params p{};
    p.M = l.M; p.N = l.N; p.K = l.K; p.n_switch = l.n_switch; p.n_off = l.n_off; p.out_scale = 1.0f / l.w_scale; p.bias = l.bias;
    p.C = l.C; p.ldc = l.ldc; p.R = l.R; p.ldr = l.ldr;
    p.Cs = l.Cs; p.cs_split_stride = l.cs_stride; p.ldcs = l.ldcs; p.relu = l.relu;
    p.overflow = ovf_ptr(h); p.stop = stop;
    const int tiles = ((l.M + tc::BM - 1) / tc::BM) * (l.N / tc::BN);
    const int grid = std::min(tiles, h->num_sms);
    // phase-stagger only when every CTA has several tiles to amortise it: a quarter of one tile's mainloop time per group
    // (~130 ns per k-block per MMA pass at the measured rates)
    if (h->opt_stagger && tiles >= 4 * grid) p.stagger_ns = (unsigned)((l.K / tc::BK) * 130 * (h->tc_fmt == 2 ? 3 : 6) / 4);
    prof_begin(h, PC_LINEAR_TC, 2.0 * l.M * (double)l.N * l.K, s);
    // TMA epilogue: only when the residual (if any) is the in-place form C += ..., which a reduce-add expresses exactly
    p.dry_store = l.dry_store;
    p.tma_out = (h->opt_tma_out && h->tc_fmt == 2 && l.Cmap && ((l.C && (!l.R || (l.R == l.C && l.ldr == l.ldc))) || (!l.C && l.Cs))) ? 1 : 0;
    const CUtensorMap& mc = l.Cmap ? *l.Cmap : *l.W;
    if (h->tc_fmt == 2 && h->opt_gemm_variant == 3 && p.tma_out) {
        auto it = h->w_half_maps.find(l.W);
        if (it != h->w_half_maps.end()) {
            const int tiles2 = ((l.M + 2 * tc2::BM - 1) / (2 * tc2::BM)) * (l.N / tc2::BN);
            const int pairs = std::max(1, std::min(tiles2, h->num_sms / 2));
            tc2::gemm2_kernel<<<2 * pairs, tc2::NUM_THREADS, tc2::SMEM_BYTES, s>>>(*l.A0, l.A1 ? *l.A1 : *l.A0, it->second, mc, p);
            prof_end(h, s);
            h->launches++;
            CU(h, cudaGetLastError());
            return FFB_OK;
        }
    }
    // small M: 128 x 64 tiles (gemm_tc.cuh Cfg<2, 1, 64>) put 4x as many CTAs on the weight stream and cut the per-tile MMA time 4x
    if (h->tc_fmt == 2 && h->opt_skinny && 2 * tiles <= h->num_sms) {
        auto it = h->w_skinny_maps.find(l.W);
        if (it != h->w_skinny_maps.end()) {
            const int tiles64 = ((l.M + tc::BM - 1) / tc::BM) * (l.N / 64);
            p.n_switch = (l.n_switch >= (1 << 28)) ? l.n_switch : l.n_switch * (tc::BN / 64);
            launch_k(h, tc::gemm_kernel<2, 1, 64>, dim3(std::min(tiles64, h->num_sms)), dim3(tc::NUM_THREADS), (size_t)tc::Cfg<2, 1, 64>::SMEM_BYTES, s,
                     *l.A0, l.A1 ? *l.A1 : *l.A0, it->second, mc, p);
            prof_end(h, s);
            h->launches++;
            CU(h, cudaGetLastError());
            return FFB_OK;
        }
    }
    // variant 1 (two staging buffers, 3 operand stages) is faster for plain / split stores, variant 0 for the in-place residual (measured,
    // profiles/tune_gemm.py --random); 2 = choose per launch
    if (h->tc_fmt == 2 && (h->opt_gemm_variant == 1 || (h->opt_gemm_variant == 2 && !l.R)))
        launch_k(h, tc::gemm_kernel<2, 1>, dim3(grid), dim3(tc::NUM_THREADS), (size_t)tc::Cfg<2, 1>::SMEM_BYTES, s, *l.A0, l.A1 ? *l.A1 : *l.A0, *l.W, mc, p);
    else if (h->tc_fmt == 2) launch_k(h, tc::gemm_kernel<2, 0>, dim3(grid), dim3(tc::NUM_THREADS), (size_t)tc::Cfg<2>::SMEM_BYTES, s, *l.A0, l.A1 ? *l.A1 : *l.A0, *l.W, mc, p);
    else launch_k(h, tc::gemm_kernel<3, 0>, dim3(grid), dim3(tc::NUM_THREADS), (size_t)tc::Cfg<3>::SMEM_BYTES, s, *l.A0, l.A1 ? *l.A1 : *l.A0, *l.W, mc, p);
    prof_end(h, s);
    h->launches++;
    CU(h, cudaGetLastError());
    return FFB_OK;
}

int launch_ln_split(ffb_handle* h, const float* x, const float* g, const float* b, uint16_t* out_plain, uint16_t* out_pos,
                    long long split_stride, const float* pos, int pos_mod, int M, int E, const int* stop, cudaStream_t s,
                    const int* pos_idx = nullptr) {
    if (M <= 0) return FFB_OK;
    prof_begin(h, PC_LAYERNORM, 8.0 * M * (double)E, s);
    if (E <= 512)
        launch_k(h, layernorm_split_kernel<4>, dim3((M + 7) / 8), dim3(256), 0, s, x, g, b, out_plain, out_pos, split_stride, pos, pos_mod, M, E, h->tc_fmt,
                 ovf_ptr(h), stop, pos_idx);
    else
        launch_k(h, layernorm_split_kernel<8>, dim3((M + 7) / 8), dim3(256), 0, s, x, g, b, out_plain, out_pos, split_stride, pos, pos_mod, M, E, h->tc_fmt,
                 ovf_ptr(h), stop, pos_idx);
    prof_end(h, s);
    h->launches++;
    CU(h, cudaGetLastError());
    return FFB_OK;
}

// Split one weight matrix [rows, K] into the operand format `fmt` (+ TMA map).  fp16x2: the matrix is pre-scaled by a
// power of two so that max|w| lands near 2^13 (both halves of the split stay in fp16's normal range); *scale_out
// receives that factor (undone in the GEMM epilogue, exact).
int split_weight(ffb_handle* h, const float* src, uint16_t* dst, size_t rows, size_t K, CUtensorMap* map, int fmt, float* scale_out,
                 cudaStream_t s) {
    float scale = 1.f;
    if (fmt == 2) {
        int* d_max = nullptr; int bits = 0;
        CU(h, cudaMalloc(&d_max, sizeof(int)));
        cudaMemsetAsync(d_max, 0, sizeof(int), s);
        absmax_kernel<<<grid1d((long long)(rows * K)), 256, 0, s>>>(src, (long long)(rows * K), d_max);
        h->launches++;
        cudaError_t e = cudaMemcpyAsync(&bits, d_max, sizeof(int), cudaMemcpyDeviceToHost, s);
        if (e == cudaSuccess) e = cudaStreamSynchronize(s);
        cudaFree(d_max);
        if (e != cudaSuccess) return fail(h, FFB_ERR_CUDA, "absmax: %s", cudaGetErrorString(e));
        float mx; memcpy(&mx, &bits, sizeof mx);
        if (!(mx <= 3.0e38f)) return fail(h, FFB_ERR_ARG, "weight matrix contains inf/nan");
        if (mx > 0.f) {
            int ex = (int)floorf(log2f(8192.0f / mx));
            ex = std::max(-40, std::min(40, ex));
            scale = ldexpf(1.0f, ex);
        }
    }
    if (scale_out) *scale_out = scale;
    const long long n4 = (long long)(rows * K / 4);
    split_array_kernel<<<grid1d(n4), 256, 0, s>>>(src, dst, n4, scale, fmt);
    h->launches++;
    CU(h, cudaGetLastError());
    return encode_weight_maps(h, map, dst, K, rows, fmt);
}

// Build (once per format) the split weights and (per batch) the activation-operand maps of format `fmt`.
int prepare_tc(ffb_handle* h, int fmt, cudaStream_t s) {
    ffb_handle::TcSet& T = h->tcs[fmt - 2];
    const size_t E = h->E, FF = h->FF, Ld = h->Ld;
    if (!T.ready) {
        const size_t per_layer = 3 * E * E + E * E + E * E + E * E + FF * E + E * FF;
        const size_t Le = h->Le, per_enc = 3 * E * E + E * E + FF * E + E * FF;
        CU(h, T.wsplit.ensure((size_t)fmt * (Ld * per_layer + E * E + Le * per_enc + 2 * Ld * E * E + E * E + E * 128) * 2));
        uint16_t* wp = T.wsplit.as<uint16_t>();
        T.layers.resize(Ld);
        for (size_t l = 0; l < Ld; ++l) {
            const DecLayerW& L = h->w.dec[l];
            ffb_handle::DecTcW& D = T.layers[l];
            D.p_sa_in = wp; FFB_TRY(split_weight(h, L.sa.in_w, wp, 3 * E, E, &D.sa_in, fmt, &D.s_sa_in, s)); wp += fmt * 3 * E * E;
            D.p_sa_out = wp; FFB_TRY(split_weight(h, L.sa.out_w, wp, E, E, &D.sa_out, fmt, &D.s_sa_out, s)); wp += fmt * E * E;
            D.p_ca_q = wp; FFB_TRY(split_weight(h, L.ca.in_w, wp, E, E, &D.ca_q, fmt, &D.s_ca_q, s)); wp += fmt * E * E;      // q rows of the cross in_proj
            D.p_ca_out = wp; FFB_TRY(split_weight(h, L.ca.out_w, wp, E, E, &D.ca_out, fmt, &D.s_ca_out, s)); wp += fmt * E * E;
            D.p_l1 = wp; FFB_TRY(split_weight(h, L.l1w, wp, FF, E, &D.l1, fmt, &D.s_l1, s)); wp += fmt * FF * E;
            D.p_l2 = wp; FFB_TRY(split_weight(h, L.l2w, wp, E, FF, &D.l2, fmt, &D.s_l2, s)); wp += fmt * E * FF;
        }
        FFB_TRY(split_weight(h, h->w.proj_w, wp, E, E, &T.proj, fmt, &T.s_proj, s)); wp += fmt * E * E;
        T.enc.resize(Le);
        for (size_t l = 0; l < Le; ++l) {
            const EncLayerW& L = h->w.enc[l];
            ffb_handle::EncTcW& D = T.enc[l];
            FFB_TRY(split_weight(h, L.sa.in_w, wp, 3 * E, E, &D.sa_in, fmt, &D.s_sa_in, s)); wp += fmt * 3 * E * E;
            FFB_TRY(split_weight(h, L.sa.out_w, wp, E, E, &D.sa_out, fmt, &D.s_sa_out, s)); wp += fmt * E * E;
            FFB_TRY(split_weight(h, L.l1w, wp, FF, E, &D.l1, fmt, &D.s_l1, s)); wp += fmt * FF * E;
            FFB_TRY(split_weight(h, L.l2w, wp, E, FF, &D.l2, fmt, &D.s_l2, s)); wp += fmt * E * FF;
        }
        FFB_TRY(split_weight(h, h->w.ckw, wp, Ld * E, E, &T.ck, fmt, &T.s_ck, s)); wp += fmt * Ld * E * E;
        FFB_TRY(split_weight(h, h->w.cvw, wp, Ld * E, E, &T.cv, fmt, &T.s_cv, s)); wp += fmt * Ld * E * E;
        FFB_TRY(split_weight(h, h->w.e2w, wp, E, E, &T.emb2, fmt, &T.s_emb2, s)); wp += fmt * E * E;
        T.has_emb0 = (h->cfg.in_dim <= 128 && h->e0pad.p != nullptr);
        if (T.has_emb0) FFB_TRY(split_weight(h, h->e0pad.as<float>(), wp, E, 128, &T.emb0, fmt, &T.s_emb0, s));
        T.ready = true;
    }
    if (h->cap_rows > 0) {
        const size_t cr = (size_t)h->cap_rows;
        FFB_TRY(encode_operand_map(h, &T.m_x2, h->a_x2.p, E, cr, tc::BM, fmt));
        FFB_TRY(encode_operand_map(h, &T.m_x2p, h->a_x2p.p, E, cr, tc::BM, fmt));
        FFB_TRY(encode_operand_map(h, &T.m_att, h->a_att.p, E, cr, tc::BM, fmt));
        FFB_TRY(encode_operand_map(h, &T.m_h, h->a_h.p, FF, cr, tc::BM, fmt));
        if (h->a_c.p) FFB_TRY(encode_operand_map(h, &T.m_c, h->a_c.p, 128, cr, tc::BM, fmt));
    }
    return FFB_OK;
}

int set_device(ffb_handle* h) {
    CU(h, cudaSetDevice(h->cfg.device));
    return FFB_OK;
}

template <class T>
int upload(ffb_handle* h, DevBuf& b, const std::vector<T>& v, cudaStream_t s) {
    CU(h, b.ensure(std::max<size_t>(v.size(), 1) * sizeof(T)));
    if (!v.empty()) CU(h, cudaMemcpyAsync(b.p, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice, s));
    return FFB_OK;
}

// ---- encode ---------------------------------------------------------------------------------------
int plan_batch(ffb_handle* h, const uint8_t* mask, const int64_t* num_input, int N, cudaStream_t s) {
    const int nl = h->cfg.num_lines, nt = h->cfg.num_token;
    h->N = N;
    h->h_row_off.assign(N + 1, 0); h->h_vlen.assign(N, 0);
    std::vector<int> nvalid(N);
    long long R = 0, Re = 0; int max_vlen = 0;
    for (int i = 0; i < N; ++i) {
        const uint8_t* m = mask + (size_t)i * nl;
        int nv = 0;
        while (nv < nl && m[nv] == 0) ++nv;
        for (int j = nv; j < nl; ++j)
            if (m[j] == 0)
                return fail(h, FFB_ERR_UNSUPPORTED, "input_mask of wireframe %d is not of prefix form (valid edge at %d after padding at %d); "
                            "only masks built like data_para.py:67-68 are supported", i, j, nv);
        nvalid[i] = nv;
        h->h_vlen[i] = nv + nt;
        h->h_row_off[i] = (int)R;
        max_vlen = std::max(max_vlen, nv + nt);
        R += nv + nt; Re += nv;
    }
    if (R > 0x7fffffffLL / std::max(h->E * 3, h->Ld * h->E)) return fail(h, FFB_ERR_ARG, "batch too large (memory rows)");
    h->h_row_off[N] = (int)R;
    h->R = R; h->Re = Re; h->max_vlen = max_vlen;
    // pos_idx / edge_src / edge_dst (one entry per memory row / edge) are expanded on the device from row_off and vlen (plan_rows_kernel below)

    std::vector<int> seq_wf, seq_first, slot_seq, seq_slot;
    h->h_seq_off.assign(N + 1, 0);
    int max_spw = 0;
    if (h->cfg.mode == FFB_MODE_PARALLEL) {
        if (!num_input) return fail(h, FFB_ERR_ARG, "num_input is required in parallel mode");
        int64_t F = 0;
        for (int i = 0; i < N; ++i) {
            if (num_input[i] < 0) return fail(h, FFB_ERR_ARG, "num_input[%d] < 0", i);
            if (num_input[i] > h->h_vlen[i])
                return fail(h, FFB_ERR_UNSUPPORTED, "num_input[%d]=%lld exceeds the %d un-masked memory rows: anchors would gather padded rows",
                            i, (long long)num_input[i], h->h_vlen[i]);
            F = std::max<int64_t>(F, num_input[i]);
        }
        if (F < 1) return fail(h, FFB_ERR_ARG, "max(num_input) must be >= 1");
        if (h->opt_force_F > 0) {
            if (h->opt_force_F < F) return fail(h, FFB_ERR_ARG, "FFB_OPT_FORCE_F = %d is smaller than this share's max(num_input) = %lld", h->opt_force_F, (long long)F);
            F = h->opt_force_F;                                     // F = max(num_input) over the whole (split) batch, model_para.py:187
        }
        h->F = (int)F;
        const bool skip_seqs = h->opt_encode_only != 0;           // encoder-only use: no sequence will be decoded, no per-sequence plan is needed
        if (!skip_seqs) slot_seq.assign((size_t)N * F, 0);
        for (int i = 0; i < N && !skip_seqs; ++i) {
            const int ni = (int)num_input[i];
            h->h_seq_off[i] = (int)seq_wf.size();
            for (int a = 0; a < ni; ++a) {                         // anchors = arange(F): rows 0..n_i-1, NOT +4 (model_para.py:201)
                slot_seq[(size_t)i * F + a] = (int)seq_wf.size();
                seq_slot.push_back((int)((size_t)i * F + a));
                seq_wf.push_back(i); seq_first.push_back(a);
            }
            if (ni < F) {                                           // padded anchors = token.len-1 (model_para.py:204-205)
                if (h->opt_dedup) {
                    int sid;
                    if (ni >= nt) sid = h->h_seq_off[i] + (nt - 1);  // identical to the real sequence anchored at row 3
                    else { sid = (int)seq_wf.size(); seq_slot.push_back((int)((size_t)i * F + ni)); seq_wf.push_back(i); seq_first.push_back(nt - 1); }
                    for (int a = ni; a < F; ++a) slot_seq[(size_t)i * F + a] = sid;
                } else {
                    for (int a = ni; a < F; ++a) {
                        slot_seq[(size_t)i * F + a] = (int)seq_wf.size();
                        seq_slot.push_back((int)((size_t)i * F + a));
                        seq_wf.push_back(i); seq_first.push_back(nt - 1);
                    }
                }
            }
            max_spw = std::max(max_spw, (int)seq_wf.size() - h->h_seq_off[i]);
        }
        h->B_full = (long long)N * F;
    } else {
        h->F = 1;
        slot_seq.resize(N);
        for (int i = 0; i < N; ++i) {
            h->h_seq_off[i] = i; slot_seq[i] = i; seq_slot.push_back(i);
            seq_wf.push_back(i); seq_first.push_back(1);           // token.SOS (model.py:190)
        }
        max_spw = 1;
        h->B_full = N;
    }
    h->h_seq_off[N] = (int)seq_wf.size();
    h->W = (h->cfg.mode == FFB_MODE_PARALLEL) ? h->opt_beam : 1;
    if (h->W > 1) {                                // beam search: hypothesis w of anchor sequence a is sequence a*W + w
        const int W = h->W;
        std::vector<int> wf2, first2, slot2;
        for (size_t a = 0; a < seq_wf.size(); ++a)
            for (int w = 0; w < W; ++w) { wf2.push_back(seq_wf[a]); first2.push_back(seq_first[a]); slot2.push_back(seq_slot[a]); }
        seq_wf.swap(wf2); seq_first.swap(first2); seq_slot.swap(slot2);
        for (auto& v : slot_seq) v *= W;           // output slot -> hypothesis 0 (the best one) of its anchor
        for (auto& v : h->h_seq_off) v *= W;
        max_spw *= W;
    }
    h->B = (long long)seq_wf.size();
    h->half_pipe = false; h->attn_x_ok = false;
    h->sum_seq_vlen = 0; h->sum_vlen2 = 0;
    for (int i = 0; i < N; ++i) {
        h->sum_seq_vlen += (double)(h->h_seq_off[i + 1] - h->h_seq_off[i]) * h->h_vlen[i];
        h->sum_vlen2 += (double)h->h_vlen[i] * h->h_vlen[i];
    }
    h->max_seq_per_wf = max_spw;
    h->encode_only = h->opt_encode_only != 0;
    const long long Mmax = h->encode_only ? 0 : h->B * (long long)(h->T - 1);
    h->pdl_on = (h->opt_pdl == 1) || (h->opt_pdl == 2 && Mmax <= 32768);      // launch-latency-bound regime
    if (Mmax * (long long)std::max(3 * h->E, h->FF) > 0x7fffffffffLL) return fail(h, FFB_ERR_ARG, "batch too large");
    if (Mmax > 0x7fffffffLL / 4) return fail(h, FFB_ERR_ARG, "batch too large (decode rows)");

    FFB_TRY(upload(h, h->d_row_off, h->h_row_off, s));
    FFB_TRY(upload(h, h->d_vlen, h->h_vlen, s));
    CU(h, h->d_pos_idx.ensure(std::max<size_t>((size_t)R, 1) * sizeof(int)));
    CU(h, h->d_edge_src.ensure(std::max<size_t>((size_t)Re, 1) * sizeof(int)));
    CU(h, h->d_edge_dst.ensure(std::max<size_t>((size_t)Re, 1) * sizeof(int)));
    plan_rows_kernel<<<grid1d(R), 256, 0, s>>>(h->d_row_off.as<int>(), N, (int)R, nl, nt, h->d_pos_idx.as<int>(), h->d_edge_src.as<int>(), h->d_edge_dst.as<int>());
    h->launches++; CU(h, cudaGetLastError());
    FFB_TRY(upload(h, h->d_seq_wf, seq_wf, s));
    FFB_TRY(upload(h, h->d_seq_first, seq_first, s));
    FFB_TRY(upload(h, h->d_seq_off, h->h_seq_off, s));
    FFB_TRY(upload(h, h->d_slot_seq, slot_seq, s));
    FFB_TRY(upload(h, h->d_seq_slot, seq_slot, s));
    {
        std::vector<int> tile_off(N + 1, 0);
        for (int i = 0; i < N; ++i) tile_off[i + 1] = tile_off[i] + (h->h_vlen[i] + 2 * al::BQ - 1) / (2 * al::BQ);     // pairs of 128-query tiles
        FFB_TRY(upload(h, h->d_tile_off, tile_off, s));
    }
    // the std::vectors above are pageable: make sure the copies are done before they go out of scope
    CU(h, cudaStreamSynchronize(s));

    // workspaces
    const size_t E = h->E, FF = h->FF, f4 = sizeof(float);
    // rows padded to the GEMM tile height: the TMA epilogue writes whole 32-row blocks (rows >= M hold don't-care values)
    const size_t rows = ((size_t)std::max<long long>(std::max<long long>(R, Mmax), 1) + 127) / 128 * 128;
    CU(h, h->mem.ensure((size_t)R * E * f4));
    const size_t Rc = h->encode_only ? 1 : (size_t)R;                       // rows of the decode-side per-wireframe caches
    const size_t Bd = h->encode_only ? 1 : (size_t)h->B;                    // sequences the decode-side buffers are sized for
    CU(h, h->Kc.ensure(Rc * h->Ld * E * f4));
    CU(h, h->Vc.ensure(Rc * h->Ld * E * f4));
    CU(h, h->tok.ensure(2 * (size_t)h->T * Bd * sizeof(int)));              // double-buffered (beam re-ordering)
    CU(h, h->beam_cum.ensure(Bd * sizeof(double)));
    h->tok_sel = 0;
    CU(h, h->logits.ensure(Bd * h->L * f4));
    CU(h, h->state.ensure(8 * sizeof(int)));
    CU(h, h->x.ensure(rows * E * f4));
    CU(h, h->x2.ensure(rows * E * f4));
    CU(h, h->qkv.ensure(rows * 3 * E * f4));
    CU(h, h->att.ensure(rows * E * f4));
    CU(h, h->hb.ensure(std::max(rows, (size_t)Re) * std::max(FF, E) * f4));
    const size_t rows_b = ((size_t)std::max<long long>((long long)Bd, 1) + 127) / 128 * 128;
    CU(h, h->xl.ensure(rows_b * E * f4));
    CU(h, h->hy32.ensure(rows_b * E * f4)); CU(h, h->memW.ensure(Rc * (E + 4) * f4));
    if (h->tc_ok && h->opt_tc) {
        h->cap_rows = (long long)((rows + 127) / 128 * 128);
        const size_t cr = (size_t)h->cap_rows;
        CU(h, h->a_x2.ensure(3 * cr * E * 2)); CU(h, h->a_x2p.ensure(3 * cr * E * 2));
        CU(h, h->a_att.ensure(3 * cr * E * 2)); CU(h, h->a_h.ensure(3 * cr * FF * 2));
        if (h->cfg.in_dim <= 128 && h->opt_enc_prec == 0) CU(h, h->a_c.ensure(2 * cr * 128 * 2));
        FFB_TRY(prepare_tc(h, h->tc_fmt, s));
        FFB_TRY(encode_split_store_map(h, &h->ms_x2, h->a_x2.p, E, cr));
        FFB_TRY(encode_output_map(h, &h->mc_x, h->x.p, E, rows));
        FFB_TRY(encode_output_map(h, &h->mc_xl, h->xl.p, E, rows_b));
        FFB_TRY(encode_output_map(h, &h->mc_qkv3, h->qkv.p, 3 * E, rows));
        FFB_TRY(encode_output_map(h, &h->mc_qkv1, h->qkv.p, E, rows));
        FFB_TRY(encode_output_map(h, &h->mc_att, h->att.p, E, rows));
        FFB_TRY(encode_split_store_map(h, &h->ms_h, h->a_h.p, FF, cr));
        h->half_pipe = (h->tc_fmt == 2 && h->opt_attn_mma == 2 && h->opt_tma_out);
        if (h->half_pipe) {
            CU(h, h->a_qkv.ensure(2 * cr * 3 * E * 2)); CU(h, h->a_qc.ensure(2 * cr * E * 2));
            CU(h, h->kc_h.ensure(2 * Rc * h->Ld * E * 2)); CU(h, h->vc_h.ensure(2 * Rc * h->Ld * E * 2));
            FFB_TRY(encode_split_store_map(h, &h->ms_qkv, h->a_qkv.p, 3 * E, cr));
            FFB_TRY(encode_split_store_map(h, &h->ms_qc, h->a_qc.p, E, cr));
            h->cap_b = (long long)rows_b;
            CU(h, h->a_ql.ensure(2 * rows_b * E * 2));
            FFB_TRY(encode_split_store_map(h, &h->ms_ql, h->a_ql.p, E, rows_b));
            h->l0_ok = false;
            if (h->opt_l0cache && !h->encode_only && h->W == 1 && h->Ld > 1) {
                CU(h, h->qkv0_cache.ensure(2 * (size_t)h->B * h->T * 3 * E * 2));
                CU(h, h->a_qkv0.ensure(2 * rows_b * 3 * E * 2));
                FFB_TRY(encode_split_store_map(h, &h->ms_qkv0, h->a_qkv0.p, 3 * E, rows_b));
                h->l0_ok = true;
            }
            h->m_last.resize(h->T);
            for (int P = 1; P < h->T && !h->encode_only; ++P)
                FFB_TRY(encode_operand_map_strided(h, &h->m_last[P], h->a_x2p.as<uint16_t>(), E, (uint64_t)h->B, (uint64_t)P, (uint64_t)(P - 1),
                                                   (uint64_t)h->cap_rows, tc::BM));
            h->attn_x_ok = (h->max_vlen <= ax::KMAX && N <= ax::MAX_GROUPS);
            if (h->attn_x_ok && !h->encode_only) {
                const size_t LdE = (size_t)h->Ld * E;
                FFB_TRY(encode_operand_map(h, &h->mx_q, h->a_qc.p, E, cr, ax::BQ, 2));
                FFB_TRY(encode_operand_map(h, &h->mx_k, h->kc_h.p, LdE, (uint64_t)R, ax::KC, 2));
                FFB_TRY(encode_rows_map(h, &h->mx_vrow, h->vc_h.p, LdE, (uint64_t)R, 64, ax::KC, CU_TENSOR_MAP_SWIZZLE_128B));
            }
            FFB_TRY(encode_output_map(h, &h->mc_kc, h->Kc.p, (uint64_t)h->Ld * E, (uint64_t)Rc));
            FFB_TRY(encode_output_map(h, &h->mc_vc, h->Vc.p, (uint64_t)h->Ld * E, (uint64_t)Rc));
            FFB_TRY(encode_rows_map(h, &h->msf_q, h->a_qkv.p, 3 * E, cr, 32, ax::BQ, CU_TENSOR_MAP_SWIZZLE_64B));
            FFB_TRY(encode_rows_map(h, &h->msf_k, h->a_qkv.p, 3 * E, cr, 32, ax::KC, CU_TENSOR_MAP_SWIZZLE_64B));
            FFB_TRY(encode_rows_map(h, &h->msf_v, h->a_qkv.p, 3 * E, cr, 64, ax::KC, CU_TENSOR_MAP_SWIZZLE_128B));
        }
    }
    return FFB_OK;
}

// Cross-attention K / V of every decoder layer, once per wireframe: k = W_k (memory + pos), v = W_v memory
// (transformer.py:248-251; torch functional.py:5866-5873).  use_tc: fp16x2 tcgen05 GEMMs (the tensor-core encoder), else fp32 SIMT.
int run_cross_cache(ffb_handle* h, cudaStream_t s, bool use_tc) {
    if (h->encode_only) return FFB_OK;
    const int E = h->E, R = (int)h->R, LdE = h->Ld * E;
    const Weights& w = h->w;
    float* mem = h->mem.as<float>();
    const int* pos_idx = h->d_pos_idx.as<int>();
    if (use_tc) {
        const ffb_handle::TcSet& TS = h->tcs[0];
        const long long ssE = h->cap_rows * E;
        split_pos_kernel<<<grid1d((long long)R * (E / 4)), 256, 0, s>>>(mem, h->a_x2.as<uint16_t>(), h->a_x2p.as<uint16_t>(), ssE, w.pos, pos_idx, R, E, 2, ovf_ptr(h));
        h->launches++;
        { TcLin l; l.A0 = &TS.m_x2p; l.W = &TS.ck; l.w_scale = TS.s_ck; l.bias = w.ckb; l.C = h->Kc.as<float>(); l.ldc = LdE; l.Cmap = &h->mc_kc;
          l.M = R; l.N = LdE; l.K = E; FFB_TRY(launch_tc(h, l, nullptr, s)); }
        { TcLin l; l.A0 = &TS.m_x2; l.W = &TS.cv; l.w_scale = TS.s_cv; l.bias = w.cvb; l.C = h->Vc.as<float>(); l.ldc = LdE;
          l.Cmap = &h->mc_vc; l.M = R; l.N = LdE; l.K = E; FFB_TRY(launch_tc(h, l, nullptr, s)); }
        return FFB_OK;
    }
    { Lin l; l.A = mem; l.lda = E; l.W = w.ckw; l.ldw = E; l.bias = w.ckb; l.C = h->Kc.as<float>(); l.ldc = LdE;
      l.pos = w.pos; l.ldpos = E; l.pos_idx = pos_idx; l.pos_cols = LdE; l.M = R; l.N = LdE; l.K = E; FFB_TRY(launch_linear(h, l, nullptr, s)); }
    { Lin l; l.A = mem; l.lda = E; l.W = w.cvw; l.ldw = E; l.bias = w.cvb; l.C = h->Vc.as<float>(); l.ldc = LdE;
      l.M = R; l.N = LdE; l.K = E; FFB_TRY(launch_linear(h, l, nullptr, s)); }
    return FFB_OK;
}

// fp16x2 copy of the cache for the half pipeline's cross-attention (overflow -> state[5], checked after the decode)
int run_cross_cache_split(ffb_handle* h, cudaStream_t s) {
    if (!h->half_pipe || h->encode_only) return FFB_OK;
    const long long n4 = (long long)h->R * h->Ld * h->E / 4;
    CU(h, cudaMemsetAsync(h->state.as<int>() + 5, 0, sizeof(int), s));
    split_array_kernel<<<grid1d(n4), 256, 0, s>>>(h->Kc.as<float>(), h->kc_h.as<uint16_t>(), n4, 1.0f, 2, h->state.as<int>() + 5);
    split_array_kernel<<<grid1d(n4), 256, 0, s>>>(h->Vc.as<float>(), h->vc_h.as<uint16_t>(), n4, 1.0f, 2, h->state.as<int>() + 5);
    h->launches += 2; CU(h, cudaGetLastError());
    return FFB_OK;
}

// ---- float64 encoder / head (enc64.cuh) -------------------------------------------------------------------
int launch_dgemm(ffb_handle* h, e64::GemmArgs a, const int* stop, cudaStream_t s) {
    if (a.M <= 0) return FFB_OK;
    a.stop = stop;
    dim3 grid((a.M + e64::GBM - 1) / e64::GBM, (a.N + e64::GBN - 1) / e64::GBN);
    prof_begin(h, PC_LINEAR, 2.0 * a.M * (double)a.N * a.K, s);
    // a single wireframe is a few hundred rows: 32 x 32 tiles put ~16x as many SMs on it (bit-identical results: same fma chain per output)
    if (2 * (int)(grid.x * grid.y) <= h->num_sms)
        e64::dgemm_kernel<32, 32, 2, 2><<<dim3((a.M + 31) / 32, (a.N + 31) / 32), 256, 0, s>>>(a);
    else
    e64::dgemm_kernel<e64::GBM, e64::GBN, 8, 4><<<grid, 256, 0, s>>>(a);
    prof_end(h, s);
    h->launches++;
    CU(h, cudaGetLastError());
    return FFB_OK;
}

int launch_ln64(ffb_handle* h, const double* x64, const float* x32, int in_mul, int in_off, const float* g, const float* b, double* y, double* yp,
                float* y32, const float* pos, const int* pos_idx, int pos_mod, int M, const int* stop, cudaStream_t s) {
    if (M <= 0) return FFB_OK;
    prof_begin(h, PC_LAYERNORM, 8.0 * M * (double)h->E, s);
    e64::layernorm64_kernel<<<(M + 7) / 8, 256, 0, s>>>(x64, x32, in_mul, in_off, g, b, y, yp, y32, pos, pos_idx, pos_mod, M, h->E, stop);
    prof_end(h, s);
    h->launches++;
    CU(h, cudaGetLastError());
    return FFB_OK;
}

// Folded pointer head: logits = memory . (W_p y + b_p) = (memory . [W_p^T ; b_p]^T) . [y ; 1]  (model_para.py:225 + 173-177).  The
// left factor depends on the wireframe only: one float64 GEMM per batch (from the float64 memory when the float64 encoder ran),
// rounded to fp32 [R, E + 4] (column E = memory . b_p).  The per-step head is then LayerNorm + one dot product per (row, sequence).
int run_head_fold(ffb_handle* h, cudaStream_t s, bool mem_is_64) {
    if (!h->opt_head64 || h->encode_only) return FFB_OK;
    e64::GemmArgs a{};
    if (mem_is_64) { a.A = h->y64.as<double>(); a.a_f32 = 0; } else { a.A = h->mem.as<float>(); a.a_f32 = 1; }
    a.lda = h->E; a.W = h->projT.as<float>(); a.ldw = h->E; a.C32 = h->memW.as<float>(); a.ldc32 = h->E + 4;
    a.M = (int)h->R; a.N = h->E + 1; a.K = h->E;
    return launch_dgemm(h, a, nullptr, s);
}

constexpr int ENC64_MAX_VLEN = 320;      // attn64_kernel keeps one row of scores per warp in (default-sized) shared memory: 8 * (320 + 64) * 8 B + the K/V tile < 48 KB

// The whole encoder and the cross-attention K / V cache in float64 (embedding.py:23-38; transformer.py:70-83,164-176,248-251).
int run_encoder64(ffb_handle* h, const float* coords_dev, cudaStream_t s) {
    const int E = h->E, FF = h->FF, R = (int)h->R, Re = (int)h->Re, N = h->N, LdE = h->Ld * E;
    const Weights& w = h->w;
    const size_t d8 = sizeof(double);
    CU(h, h->x64.ensure((size_t)R * E * d8)); CU(h, h->y64.ensure((size_t)R * E * d8)); CU(h, h->yp64.ensure((size_t)R * E * d8));
    CU(h, h->qkv64.ensure((size_t)R * 3 * E * d8)); CU(h, h->att64.ensure((size_t)R * E * d8));
    CU(h, h->h64.ensure((size_t)std::max(R, Re) * std::max(FF, E) * d8));
    double* x = h->x64.as<double>(); double* y = h->y64.as<double>(); double* yp = h->yp64.as<double>();
    double* qkv = h->qkv64.as<double>(); double* att = h->att64.as<double>(); double* hb = h->h64.as<double>();
    const int* row_off = h->d_row_off.as<int>(); const int* vlen = h->d_vlen.as<int>(); const int* pos_idx = h->d_pos_idx.as<int>();
    // value embedding on the valid edges, token rows
    { e64::GemmArgs a{}; a.A = coords_dev; a.lda = h->cfg.in_dim; a.a_f32 = 1; a.a_rows = h->d_edge_src.as<int>(); a.W = w.e0w; a.ldw = h->cfg.in_dim;
      a.bias = w.e0b; a.C = hb; a.ldc = E; a.M = Re; a.N = E; a.K = h->cfg.in_dim; a.relu = 1; FFB_TRY(launch_dgemm(h, a, nullptr, s)); }
    { e64::GemmArgs a{}; a.A = hb; a.lda = E; a.W = w.e2w; a.ldw = E; a.bias = w.e2b; a.C = x; a.ldc = E; a.c_rows = h->d_edge_dst.as<int>();
      a.M = Re; a.N = E; a.K = E; FFB_TRY(launch_dgemm(h, a, nullptr, s)); }
    e64::token_rows64_kernel<<<grid1d((long long)N * h->cfg.num_token * E), 256, 0, s>>>(w.tok_table, row_off, x, N, h->cfg.num_token, E);
    h->launches++; CU(h, cudaGetLastError());
    const size_t a_smem = e64::A64_ROWS * (size_t)(h->max_vlen + 64) * d8;
    for (int li = 0; li < h->Le; ++li) {                                  // TransformerEncoderLayer.forward_pre (transformer.py:164-176)
        const EncLayerW& L = w.enc[li];
        FFB_TRY(launch_ln64(h, x, nullptr, 1, 0, L.n1w, L.n1b, y, yp, nullptr, w.pos, pos_idx, 1, R, nullptr, s));     // q = k = LN(x) + pos, v = LN(x)
        { e64::GemmArgs a{}; a.A = yp; a.lda = E; a.W = L.sa.in_w; a.ldw = E; a.bias = L.sa.in_b; a.C = qkv; a.ldc = 3 * E; a.M = R; a.N = 2 * E; a.K = E;
          FFB_TRY(launch_dgemm(h, a, nullptr, s)); }
        { e64::GemmArgs a{}; a.A = y; a.lda = E; a.W = L.sa.in_w + (size_t)2 * E * E; a.ldw = E; a.bias = L.sa.in_b + 2 * E; a.C = qkv + 2 * E; a.ldc = 3 * E;
          a.M = R; a.N = E; a.K = E; FFB_TRY(launch_dgemm(h, a, nullptr, s)); }
        prof_begin(h, PC_ATTN_TILED, 4.0 * 64 * h->H * h->sum_vlen2, s);
        e64::attn64_kernel<<<dim3((h->max_vlen + e64::A64_ROWS - 1) / e64::A64_ROWS, h->H, N), 32 * e64::A64_ROWS, a_smem, s>>>(qkv, att, row_off, vlen, E, h->max_vlen);
        prof_end(h, s);
        h->launches++; CU(h, cudaGetLastError());
        { e64::GemmArgs a{}; a.A = att; a.lda = E; a.W = L.sa.out_w; a.ldw = E; a.bias = L.sa.out_b; a.C = x; a.ldc = E; a.R = x; a.ldr = E;
          a.M = R; a.N = E; a.K = E; FFB_TRY(launch_dgemm(h, a, nullptr, s)); }
        FFB_TRY(launch_ln64(h, x, nullptr, 1, 0, L.n2w, L.n2b, y, nullptr, nullptr, nullptr, nullptr, 1, R, nullptr, s));
        { e64::GemmArgs a{}; a.A = y; a.lda = E; a.W = L.l1w; a.ldw = E; a.bias = L.l1b; a.C = hb; a.ldc = FF; a.M = R; a.N = FF; a.K = E; a.relu = 1;
          FFB_TRY(launch_dgemm(h, a, nullptr, s)); }
        { e64::GemmArgs a{}; a.A = hb; a.lda = FF; a.W = L.l2w; a.ldw = FF; a.bias = L.l2b; a.C = x; a.ldc = E; a.R = x; a.ldr = E;
          a.M = R; a.N = E; a.K = FF; FFB_TRY(launch_dgemm(h, a, nullptr, s)); }
    }
    // encoder.norm (transformer.py:80-81): memory in float64 (y), memory + pos (yp) and its fp32 rounding (what the decode gathers and scores)
    FFB_TRY(launch_ln64(h, x, nullptr, 1, 0, w.enc_nw, w.enc_nb, y, yp, h->mem.as<float>(), w.pos, pos_idx, 1, R, nullptr, s));
    // cross-attention K / V of every decoder layer.  Large batches: fp16x2 tcgen05 GEMMs on the fp32 rounding of the float64 memory (the
    // products are exact to 2^-22, the class of the decoder that consumes them; 24 of the encoder's 135 GFLOP on the bench batch leave the
    // FP64 pipe).  Otherwise from the float64 memory on the FP64 pipe, rounded to fp32 once.
    const ffb_handle::TcSet& TS = h->tcs[0];
    if (h->half_pipe && h->tc_fmt == 2 && TS.ready && R >= TC_MIN_ROWS && !h->encode_only) {
        FFB_TRY(run_cross_cache(h, s, true));
    } else if (!h->encode_only) {
        { e64::GemmArgs a{}; a.A = yp; a.lda = E; a.W = w.ckw; a.ldw = E; a.bias = w.ckb; a.C32 = h->Kc.as<float>(); a.ldc32 = LdE; a.M = R; a.N = LdE; a.K = E;
          FFB_TRY(launch_dgemm(h, a, nullptr, s)); }
        { e64::GemmArgs a{}; a.A = y; a.lda = E; a.W = w.cvw; a.ldw = E; a.bias = w.cvb; a.C32 = h->Vc.as<float>(); a.ldc32 = LdE; a.M = R; a.N = LdE; a.K = E;
          FFB_TRY(launch_dgemm(h, a, nullptr, s)); }
    }
    FFB_TRY(run_head_fold(h, s, true));
    return run_cross_cache_split(h, s);
}

int run_encoder(ffb_handle* h, const float* coords_dev, cudaStream_t s, bool allow_tc) {
    const int E = h->E, FF = h->FF, R = (int)h->R, Re = (int)h->Re, N = h->N;
    const Weights& w = h->w;
    float* x = h->x.as<float>(); float* x2 = h->x2.as<float>(); float* qkv = h->qkv.as<float>();
    float* att = h->att.as<float>(); float* hb = h->hb.as<float>(); float* mem = h->mem.as<float>();
    const int* row_off = h->d_row_off.as<int>(); const int* vlen = h->d_vlen.as<int>();
    const int* pos_idx = h->d_pos_idx.as<int>();

    // Encoder layers + cross K / V projections on the tcgen05 pipeline (fp16x2 GEMMs, tcgen05 attention) when the batch is large
    // enough; otherwise fp32 SIMT GEMMs + the 3xTF32 mma.sync attention kernel.
    const ffb_handle::TcSet& TS = h->tcs[0];
    const bool enc_tc = allow_tc && h->opt_enc_tc && h->half_pipe && h->tc_fmt == 2 && TS.ready &&
                        (int)TS.enc.size() == h->Le && (h->opt_tc == 2 || R >= TC_MIN_ROWS);
    // value embedding (embedding.py:34): relu(coords W0^T + b0) W2^T + b2 on the valid edges only
    if (enc_tc) {
        // both linears run over ALL memory rows on the tcgen05 GEMM (the 4 token rows per wireframe compute don't-care values that
        // token_rows_kernel overwrites); the first one needs K = in_dim (100) zero-padded to 128, else it stays on the FFMA kernel
        h->ovf_slot = 6;
        CU(h, cudaMemsetAsync(h->state.as<int>() + 6, 0, sizeof(int), s));
        if (TS.has_emb0 && h->a_c.p) {
            // both linears on tcgen05: the coordinates become an fp16x2 operand with K zero-padded to 128, written at the edges' memory rows
            split_coords_kernel<<<grid1d((long long)Re * 32), 256, 0, s>>>(coords_dev, h->d_edge_src.as<int>(), h->d_edge_dst.as<int>(), h->a_c.as<uint16_t>(),
                                                                        h->cap_rows * 128, Re, h->cfg.in_dim, ovf_ptr(h));
            h->launches++; CU(h, cudaGetLastError());
            { TcLin l; l.A0 = &TS.m_c; l.W = &TS.emb0; l.w_scale = TS.s_emb0; l.bias = w.e0b; l.relu = 1; l.Cs = h->a_x2.as<uint16_t>();
              l.cs_stride = h->cap_rows * E; l.ldcs = E; l.Cmap = &h->ms_x2; l.M = R; l.N = E; l.K = 128; FFB_TRY(launch_tc(h, l, nullptr, s)); }
        } else {
            CU(h, cudaMemsetAsync(hb, 0, (size_t)R * E * sizeof(float), s));
            { Lin l; l.A = coords_dev; l.lda = h->cfg.in_dim; l.a_rows = h->d_edge_src.as<int>(); l.W = w.e0w; l.ldw = h->cfg.in_dim; l.bias = w.e0b;
              l.C = hb; l.ldc = E; l.c_rows = h->d_edge_dst.as<int>(); l.M = Re; l.N = E; l.K = h->cfg.in_dim; l.relu = 1; FFB_TRY(launch_linear(h, l, nullptr, s)); }
            split_rows_kernel<<<grid1d((long long)R * (E / 4)), 256, 0, s>>>(hb, h->a_x2.as<uint16_t>(), h->cap_rows * E, R, E, 2, ovf_ptr(h));
            h->launches++; CU(h, cudaGetLastError());
        }
        { TcLin l; l.A0 = &TS.m_x2; l.W = &TS.emb2; l.w_scale = TS.s_emb2; l.bias = w.e2b; l.C = x; l.ldc = E; l.Cmap = &h->mc_x; l.M = R; l.N = E; l.K = E;
          FFB_TRY(launch_tc(h, l, nullptr, s)); }
        h->ovf_slot = 4;
    } else {
        { Lin l; l.A = coords_dev; l.lda = h->cfg.in_dim; l.a_rows = h->d_edge_src.as<int>(); l.W = w.e0w; l.ldw = h->cfg.in_dim; l.bias = w.e0b;
          l.C = hb; l.ldc = E; l.M = Re; l.N = E; l.K = h->cfg.in_dim; l.relu = 1; FFB_TRY(launch_linear(h, l, nullptr, s)); }
        { Lin l; l.A = hb; l.lda = E; l.W = w.e2w; l.ldw = E; l.bias = w.e2b; l.C = x; l.ldc = E; l.c_rows = h->d_edge_dst.as<int>();
          l.M = Re; l.N = E; l.K = E; FFB_TRY(launch_linear(h, l, nullptr, s)); }
    }
    token_rows_kernel<<<grid1d((long long)N * h->cfg.num_token * (E / 4)), 256, 0, s>>>(w.tok_table, row_off, x, N, h->cfg.num_token, E);
    h->launches++; CU(h, cudaGetLastError());

    const bool enc_ax = h->attn_x_ok && (h->opt_attn_x & 1);             // tcgen05 attention needs <= 256 rows per wireframe
    h->enc_used_tc = enc_tc;
    const int LdE = h->Ld * E;
    if (enc_tc) {
        uint16_t* ax2 = h->a_x2.as<uint16_t>(); uint16_t* ax2p = h->a_x2p.as<uint16_t>();
        uint16_t* aatt = h->a_att.as<uint16_t>(); uint16_t* ah = h->a_h.as<uint16_t>(); uint16_t* aqkv = h->a_qkv.as<uint16_t>();
        const long long ssE = h->cap_rows * E, ssF = h->cap_rows * FF;
        h->ovf_slot = 6;                                                  // an fp16-range overflow here makes ffb_encode redo the encoder in fp32
        int rc = FFB_OK;                                                  // (state[6] was cleared before the embedding above)
        for (int li = 0; li < h->Le && rc == FFB_OK; ++li) {              // TransformerEncoderLayer.forward_pre (transformer.py:164-176)
            const EncLayerW& L = w.enc[li];
            const ffb_handle::EncTcW& Tw = TS.enc[li];
            rc = launch_ln_split(h, x, L.n1w, L.n1b, ax2, ax2p, ssE, w.pos, 1, R, E, nullptr, s, pos_idx);   // q = k = LN(x) + pos, v = LN(x)
            if (rc != FFB_OK) break;
            { TcLin l; l.A0 = &TS.m_x2p; l.A1 = &TS.m_x2; l.n_switch = 2 * E / tc::BN; l.W = &Tw.sa_in; l.w_scale = Tw.s_sa_in; l.bias = L.sa.in_b;
              l.M = R; l.N = 3 * E; l.K = E; l.Cs = aqkv; l.cs_stride = h->cap_rows * 3 * E; l.ldcs = 3 * E; l.Cmap = &h->ms_qkv;
              if ((rc = launch_tc(h, l, nullptr, s)) != FFB_OK) break; }
            if (enc_ax) {
              ax::Params ap{}; ap.mode = 0; ap.seq_off = row_off; ap.q_mul = 1; ap.row_off = row_off; ap.vlen = vlen; ap.n_groups = N;
              ap.q_col = 0; ap.k_col = E; ap.v_col = 2 * E; ap.Os = aatt; ap.os_stride = ssE; ap.ldo = E;
              if ((rc = launch_attn_x(h, h->msf_q, h->msf_k, h->msf_v, ap, &h->h_row_off, h->sum_vlen2, PC_ATTN_TILED, nullptr, s)) != FFB_OK) break;
            } else if (h->opt_attn_long) {                                // > 256 keys per wireframe: K / V stream through the tcgen05 kernel
              al::Params lp{}; lp.tile_off = h->d_tile_off.as<int>(); lp.row_off = row_off; lp.vlen = vlen; lp.n_groups = N;
              lp.q_col = 0; lp.k_col = E; lp.v_col = 2 * E; lp.Os = aatt; lp.os_stride = ssE; lp.ldo = E;
              long long tiles = 0;
              for (int i = 0; i < N; ++i) tiles += (h->h_vlen[i] + 2 * al::BQ - 1) / (2 * al::BQ);
              if ((rc = launch_attn_long(h, h->msf_q, h->msf_k, h->msf_v, lp, tiles, h->sum_vlen2, PC_ATTN_TILED, s)) != FFB_OK) break;
            } else {                                                      // fp16x2 mma.sync kernel on the same operands
              AttnHalfIn in{aqkv, h->cap_rows * 3 * E, 3 * E, aqkv + E, h->cap_rows * 3 * E, aqkv + 2 * E, h->cap_rows * 3 * E, 3 * E};
              AttnGroups g{}; g.ragged = 1; g.q_begin = row_off; g.q_mul = 1; g.k_begin = row_off; g.k_len = vlen;
              if ((rc = launch_attn_h(h, in, aatt, ssE, g, N, h->max_vlen, h->max_vlen, h->sum_vlen2, PC_ATTN_TILED, nullptr, s)) != FFB_OK) break;
            }
            { TcLin l; l.A0 = &TS.m_att; l.W = &Tw.sa_out; l.w_scale = Tw.s_sa_out; l.bias = L.sa.out_b; l.C = x; l.ldc = E; l.R = x; l.ldr = E;
              l.Cmap = &h->mc_x; l.M = R; l.N = E; l.K = E; if ((rc = launch_tc(h, l, nullptr, s)) != FFB_OK) break; }
            if ((rc = launch_ln_split(h, x, L.n2w, L.n2b, ax2, nullptr, ssE, nullptr, 1, R, E, nullptr, s)) != FFB_OK) break;
            { TcLin l; l.A0 = &TS.m_x2; l.W = &Tw.l1; l.w_scale = Tw.s_l1; l.bias = L.l1b; l.relu = 1; l.Cs = ah; l.cs_stride = ssF; l.ldcs = FF;
              l.Cmap = &h->ms_h; l.M = R; l.N = FF; l.K = E; if ((rc = launch_tc(h, l, nullptr, s)) != FFB_OK) break; }
            { TcLin l; l.A0 = &TS.m_h; l.W = &Tw.l2; l.w_scale = Tw.s_l2; l.bias = L.l2b; l.C = x; l.ldc = E; l.R = x; l.ldr = E;
              l.Cmap = &h->mc_x; l.M = R; l.N = E; l.K = FF; if ((rc = launch_tc(h, l, nullptr, s)) != FFB_OK) break; }
        }
        if (rc == FFB_OK) rc = launch_ln(h, x, w.enc_nw, w.enc_nb, mem, R, E, nullptr, s);   // encoder.norm (transformer.py:80-81)
        if (rc == FFB_OK) rc = run_cross_cache(h, s, true);
        h->ovf_slot = 4;
        FFB_TRY(rc);
        CU(h, cudaGetLastError());
    } else {
    AttnGroups g{}; g.ragged = 1; g.q_begin = row_off; g.q_mul = 1; g.k_begin = row_off; g.k_len = vlen;
    for (int li = 0; li < h->Le; ++li) {                                  // TransformerEncoderLayer.forward_pre (transformer.py:164-176)
        const EncLayerW& L = w.enc[li];
        FFB_TRY(launch_ln(h, x, L.n1w, L.n1b, x2, R, E, nullptr, s));
        { Lin l; l.A = x2; l.lda = E; l.W = L.sa.in_w; l.ldw = E; l.bias = L.sa.in_b; l.C = qkv; l.ldc = 3 * E;
          l.pos = w.pos; l.ldpos = E; l.pos_idx = pos_idx; l.pos_cols = 2 * E; l.M = R; l.N = 3 * E; l.K = E;
          FFB_TRY(launch_linear(h, l, nullptr, s)); }
        FFB_TRY(launch_attn_tiled(h, qkv, 3 * E, qkv + E, qkv + 2 * E, 3 * E, att, E, g, N, h->max_vlen, h->sum_vlen2, nullptr, s));
        { Lin l; l.A = att; l.lda = E; l.W = L.sa.out_w; l.ldw = E; l.bias = L.sa.out_b; l.C = x; l.ldc = E; l.R = x; l.ldr = E;
          l.M = R; l.N = E; l.K = E; FFB_TRY(launch_linear(h, l, nullptr, s)); }
        FFB_TRY(launch_ln(h, x, L.n2w, L.n2b, x2, R, E, nullptr, s));
        { Lin l; l.A = x2; l.lda = E; l.W = L.l1w; l.ldw = E; l.bias = L.l1b; l.C = hb; l.ldc = FF; l.M = R; l.N = FF; l.K = E; l.relu = 1;
          FFB_TRY(launch_linear(h, l, nullptr, s)); }
        { Lin l; l.A = hb; l.lda = FF; l.W = L.l2w; l.ldw = FF; l.bias = L.l2b; l.C = x; l.ldc = E; l.R = x; l.ldr = E;
          l.M = R; l.N = E; l.K = FF; FFB_TRY(launch_linear(h, l, nullptr, s)); }
    }
    FFB_TRY(launch_ln(h, x, w.enc_nw, w.enc_nb, mem, R, E, nullptr, s));   // encoder.norm (transformer.py:80-81)
    FFB_TRY(run_cross_cache(h, s, false));
    }
    FFB_TRY(run_head_fold(h, s, false));
    return run_cross_cache_split(h, s);
}

// ---- one decode step --------------------------------------------------------------------------------
// Runs the loop body for prefix length P on tok[0..P) and (if tok_out) appends tok[P].
int run_step(ffb_handle* h, int P, bool append, cudaStream_t s) {
    const int E = h->E, FF = h->FF, B = (int)h->B, N = h->N, Ld = h->Ld;
    const int M = B * P;
    const Weights& w = h->w;
    float* x = h->x.as<float>(); float* x2 = h->x2.as<float>(); float* qkv = h->qkv.as<float>();
    float* att = h->att.as<float>(); float* hb = h->hb.as<float>(); float* xl = h->xl.as<float>();
    int* st = h->state.as<int>();
    const int* stop = st;            // state[0]
    const int* row_off = h->d_row_off.as<int>(); const int* vlen = h->d_vlen.as<int>();
    const int* seq_off = h->d_seq_off.as<int>();
    const int LdE = Ld * E;

    prof_begin(h, PC_OTHER, 0.0, s);
    launch_k(h, gather_tgt_kernel, dim3(grid1d((long long)M * (E / 4))), dim3(256), 0, s, (const float*)h->mem.as<float>(), row_off,
             (const int*)h->d_seq_wf.as<int>(), (const int*)tok_cur(h), x, B, P, E, stop);
    prof_end(h, s);
    h->launches++; CU(h, cudaGetLastError());

    float* cur = x; int rows = M; int Pq = P;            // rows carried through the rest of the layer
    const ffb_handle::TcSet& TS = h->tcs[h->tc_fmt - 2];
    const bool tc = h->tc_ok && h->opt_tc && TS.ready && (int)TS.layers.size() == Ld && (h->opt_tc == 2 || M >= TC_MIN_ROWS_DECODE);
    uint16_t* ax2 = h->a_x2.as<uint16_t>(); uint16_t* ax2p = h->a_x2p.as<uint16_t>();
    uint16_t* aatt = h->a_att.as<uint16_t>(); uint16_t* ah = h->a_h.as<uint16_t>();
    uint16_t* aqkv = h->a_qkv.as<uint16_t>(); uint16_t* aqc = h->a_qc.as<uint16_t>();
    // teacher-forced pass (ffb_forward_train): causal + label-padding masks in the self-attention; they exist in the fp32 attention kernel
    // only, so q,k,v stay fp32 (no half pipeline) while the GEMMs remain on the tensor path
    const unsigned char* tmask = h->train_kmask;
    const int causal = tmask ? 1 : 0;
    const bool hp = h->half_pipe && h->tc_fmt == 2 && !tmask;      // q,k,v / cross-q produced and consumed as fp16x2 (no fp32 copy)
    const long long ssE = h->cap_rows * E, ssF = h->cap_rows * FF;     // elements between the operand splits
    for (int li = 0; li < Ld; ++li) {                    // TransformerDecoderLayer.forward_pre (transformer.py:235-256)
        const DecLayerW& Lw = w.dec[li];
        const bool last = h->opt_prune && (li == Ld - 1);
        const float* qpos_cross = last ? w.qpos + (size_t)(P - 1) * E : w.qpos;
        const int qmod_cross = last ? 1 : P;
        if (!tc) {
            // ---- fp32 SIMT path (small M, or geometries the tensor-core kernel does not cover) ----
            // self-attention over the whole prefix, NO causal mask (model_para.py:222-223)
            FFB_TRY(launch_ln(h, x, Lw.n1w, Lw.n1b, x2, M, E, stop, s));
            { Lin l; l.A = x2; l.lda = E; l.W = Lw.sa.in_w; l.ldw = E; l.bias = Lw.sa.in_b; l.C = qkv; l.ldc = 3 * E;
              l.pos = w.qpos; l.ldpos = E; l.pos_mod = P; l.pos_cols = 2 * E; l.M = M; l.N = 3 * E; l.K = E;
              FFB_TRY(launch_linear(h, l, stop, s)); }
            if (!last) {
                FFB_TRY(launch_attn_rows(h, qkv, 3 * E, qkv + E, qkv + 2 * E, 3 * E, att, E, B, P, P, P, 0, P, P, stop, s, nullptr, 0, tmask, causal));
                { Lin l; l.A = att; l.lda = E; l.W = Lw.sa.out_w; l.ldw = E; l.bias = Lw.sa.out_b; l.C = x; l.ldc = E; l.R = x; l.ldr = E;
                  l.M = M; l.N = E; l.K = E; FFB_TRY(launch_linear(h, l, stop, s)); }
            } else {
                // only pointer[-1] is consumed (model_para.py:176): carry just the last position from here on
                FFB_TRY(launch_attn_rows(h, qkv, 3 * E, qkv + E, qkv + 2 * E, 3 * E, att, E, B, 1, P, P, P - 1, P, 1, stop, s));
                launch_k(h, copy_rows_kernel, dim3(grid1d((long long)B * (E / 4))), dim3(256), 0, s, (const float*)x, xl, B, P, P - 1, E, stop);
                h->launches++; CU(h, cudaGetLastError());
                { Lin l; l.A = att; l.lda = E; l.W = Lw.sa.out_w; l.ldw = E; l.bias = Lw.sa.out_b; l.C = xl; l.ldc = E; l.R = xl; l.ldr = E;
                  l.M = B; l.N = E; l.K = E; FFB_TRY(launch_linear(h, l, stop, s)); }
                cur = xl; rows = B; Pq = 1;
            }
            // cross-attention over the cached K/V of the owning wireframe
            FFB_TRY(launch_ln(h, cur, Lw.n2w, Lw.n2b, x2, rows, E, stop, s));
            { Lin l; l.A = x2; l.lda = E; l.W = Lw.ca.in_w; l.ldw = E; l.bias = Lw.ca.in_b; l.C = qkv; l.ldc = E;
              l.pos = qpos_cross; l.ldpos = E; l.pos_mod = qmod_cross; l.pos_cols = E; l.M = rows; l.N = E; l.K = E;
              FFB_TRY(launch_linear(h, l, stop, s)); }
            { AttnGroups g{}; g.ragged = 1; g.q_begin = seq_off; g.q_mul = Pq; g.k_begin = row_off; g.k_len = vlen;
              FFB_TRY(launch_attn_tiled(h, qkv, E, h->Kc.as<float>() + (size_t)li * E, h->Vc.as<float>() + (size_t)li * E, LdE,
                                        att, E, g, N, h->max_seq_per_wf * Pq, h->sum_seq_vlen * Pq, stop, s)); }
            { Lin l; l.A = att; l.lda = E; l.W = Lw.ca.out_w; l.ldw = E; l.bias = Lw.ca.out_b; l.C = cur; l.ldc = E; l.R = cur; l.ldr = E;
              l.M = rows; l.N = E; l.K = E; FFB_TRY(launch_linear(h, l, stop, s)); }
            // feed-forward
            FFB_TRY(launch_ln(h, cur, Lw.n3w, Lw.n3b, x2, rows, E, stop, s));
            { Lin l; l.A = x2; l.lda = E; l.W = Lw.l1w; l.ldw = E; l.bias = Lw.l1b; l.C = hb; l.ldc = FF; l.M = rows; l.N = FF; l.K = E; l.relu = 1;
              FFB_TRY(launch_linear(h, l, stop, s)); }
            { Lin l; l.A = hb; l.lda = FF; l.W = Lw.l2w; l.ldw = FF; l.bias = Lw.l2b; l.C = cur; l.ldc = E; l.R = cur; l.ldr = E;
              l.M = rows; l.N = E; l.K = FF; FFB_TRY(launch_linear(h, l, stop, s)); }
        } else {
            // ---- tensor-core path: every GEMM operand is produced directly in the split operand format (fp16x2 / bf16x3) ----
            const ffb_handle::DecTcW& Tw = TS.layers[li];
            // layer 0 inside the greedy loop: q / k / v of the positions that existed in the previous step are unchanged (exact): project the
            // NEW position only (B rows instead of B * P) and assemble the (sequence, position)-ordered operand from the cache
            const bool l0 = (li == 0) && append && hp && h->l0_ok && !h->l0_suspend && !last;
            if (l0) {
                launch_k(h, copy_rows_kernel, dim3(grid1d((long long)B * (E / 4))), dim3(256), 0, s, (const float*)x, xl, B, P, P - 1, E, stop);
                h->launches++; CU(h, cudaGetLastError());
                FFB_TRY(launch_ln_split(h, xl, Lw.n1w, Lw.n1b, ax2, ax2p, ssE, w.qpos + (size_t)(P - 1) * E, 1, B, E, stop, s));
                { TcLin l; l.A0 = &TS.m_x2p; l.A1 = &TS.m_x2; l.n_switch = 2 * E / tc::BN; l.W = &Tw.sa_in; l.w_scale = Tw.s_sa_in; l.bias = Lw.sa.in_b;
                  l.M = B; l.N = 3 * E; l.K = E; l.Cs = h->a_qkv0.as<uint16_t>(); l.cs_stride = h->cap_b * 3 * E; l.ldcs = 3 * E; l.Cmap = &h->ms_qkv0;
                  FFB_TRY(launch_tc(h, l, stop, s)); }
                prof_begin(h, PC_OTHER, 0.0, s);
                launch_k(h, assemble_qkv0_kernel, dim3(grid1d(2ll * M * (3 * E / 8))), dim3(256), 0, s, (const uint16_t*)h->a_qkv0.as<uint16_t>(), h->cap_b * 3 * E,
                         h->qkv0_cache.as<uint16_t>(), (long long)h->B * h->T * 3 * E, aqkv, h->cap_rows * 3 * E, B, P, h->T, 3 * E, stop);
                prof_end(h, s);
                h->launches++; CU(h, cudaGetLastError());
            } else
            FFB_TRY(launch_ln_split(h, x, Lw.n1w, Lw.n1b, ax2, ax2p, ssE, w.qpos, P, M, E, stop, s));
            const bool q_last_only = last && hp && h->H <= 8;        // pruned last layer: q is needed for the last prefix position only
            if (l0) {
            } else if (q_last_only) {
                // k, v for every position: the column window [E, 3E) of the in-projection; q for the rows b*P + P-1 only, through a
                // strided view of the LayerNorm output, into the compact buffer a_ql
                { TcLin l; l.A0 = &TS.m_x2p; l.A1 = &TS.m_x2; l.n_switch = E / tc::BN; l.n_off = E;
                  l.W = &Tw.sa_in; l.w_scale = Tw.s_sa_in; l.bias = Lw.sa.in_b; l.M = M; l.N = 2 * E; l.K = E;
                  l.Cs = aqkv; l.cs_stride = h->cap_rows * 3 * E; l.ldcs = 3 * E; l.Cmap = &h->ms_qkv;
                  FFB_TRY(launch_tc(h, l, stop, s)); }
                { TcLin l; l.A0 = &h->m_last[P]; /* encoded once per batch (plan_batch) */ l.W = &Tw.sa_in; l.w_scale = Tw.s_sa_in; l.bias = Lw.sa.in_b; l.M = B; l.N = E; l.K = E;
                  l.Cs = h->a_ql.as<uint16_t>(); l.cs_stride = h->cap_b * E; l.ldcs = E; l.Cmap = &h->ms_ql;
                  FFB_TRY(launch_tc(h, l, stop, s)); }
            } else
            { TcLin l; l.A0 = &TS.m_x2p; l.A1 = &TS.m_x2; l.n_switch = 2 * E / tc::BN;      // q,k from x2+qpos; v from x2
              l.W = &Tw.sa_in; l.w_scale = Tw.s_sa_in; l.bias = Lw.sa.in_b; l.M = M; l.N = 3 * E; l.K = E;
              if (hp) { l.Cs = aqkv; l.cs_stride = h->cap_rows * 3 * E; l.ldcs = 3 * E; l.Cmap = &h->ms_qkv; }   // q,k,v straight to fp16x2
              else { l.C = qkv; l.ldc = 3 * E; l.Cmap = &h->mc_qkv3; }
              FFB_TRY(launch_tc(h, l, stop, s)); }
            if (!last) {
                if (hp && (h->opt_attn_x & 2) && P <= ax::BQ) {
                    ax::Params ap{}; ap.mode = 1; ap.P = P; ap.n_seqs = B; ap.q_col = 0; ap.k_col = E; ap.v_col = 2 * E;
                    ap.Os = aatt; ap.os_stride = ssE; ap.ldo = E;
                    FFB_TRY(launch_attn_x(h, h->msf_q, h->msf_k, h->msf_v, ap, nullptr, (double)B * P * P, PC_ATTN_ROWS, stop, s));
                } else if (hp) {
                    AttnHalfIn in{aqkv, h->cap_rows * 3 * E, 3 * E, aqkv + E, h->cap_rows * 3 * E, aqkv + 2 * E, h->cap_rows * 3 * E, 3 * E};
                    AttnGroups g{}; g.ragged = 0; g.nq = P; g.nk = P; g.q_stride = P; g.q_off = 0; g.k_stride = P; g.o_stride = P;
                    FFB_TRY(launch_attn_h(h, in, aatt, ssE, g, B, P, P, (double)B * P * P, PC_ATTN_ROWS, stop, s));
                } else
                FFB_TRY(launch_attn_rows(h, qkv, 3 * E, qkv + E, qkv + 2 * E, 3 * E, att, E, B, P, P, P, 0, P, P, stop, s, aatt, ssE, tmask, causal));
                { TcLin l; l.A0 = &TS.m_att; l.W = &Tw.sa_out; l.w_scale = Tw.s_sa_out; l.bias = Lw.sa.out_b; l.C = x; l.ldc = E; l.R = x; l.ldr = E; l.Cmap = &h->mc_x;
                  l.M = M; l.N = E; l.K = E; FFB_TRY(launch_tc(h, l, stop, s)); }
            } else {
                if (hp && h->H <= 8) {       // one CTA per sequence, one warp per head (attn_h.cuh: attn_last_kernel)
                    prof_begin(h, PC_ATTN_ROWS, 4.0 * 64 * h->H * (double)B * P, s);
                    launch_k(h, attn_last_kernel, dim3(B), dim3(32 * h->H), 0, s, (const uint16_t*)aqkv, h->cap_rows * 3 * E, 3 * E, E, P, B, aatt, ssE, E, stop,
                             (const uint16_t*)h->a_ql.as<uint16_t>(), h->cap_b * E);
                    prof_end(h, s);
                    h->launches++; CU(h, cudaGetLastError());
                } else if (hp) {
                    AttnHalfIn in{aqkv, h->cap_rows * 3 * E, 3 * E, aqkv + E, h->cap_rows * 3 * E, aqkv + 2 * E, h->cap_rows * 3 * E, 3 * E};
                    AttnGroups g{}; g.ragged = 0; g.nq = 1; g.nk = P; g.q_stride = P; g.q_off = P - 1; g.k_stride = P; g.o_stride = 1;
                    FFB_TRY(launch_attn_h(h, in, aatt, ssE, g, B, 1, P, (double)B * P, PC_ATTN_ROWS, stop, s));
                } else
                FFB_TRY(launch_attn_rows(h, qkv, 3 * E, qkv + E, qkv + 2 * E, 3 * E, att, E, B, 1, P, P, P - 1, P, 1, stop, s, aatt, ssE));
                launch_k(h, copy_rows_kernel, dim3(grid1d((long long)B * (E / 4))), dim3(256), 0, s, (const float*)x, xl, B, P, P - 1, E, stop);
                h->launches++; CU(h, cudaGetLastError());
                { TcLin l; l.A0 = &TS.m_att; l.W = &Tw.sa_out; l.w_scale = Tw.s_sa_out; l.bias = Lw.sa.out_b; l.C = xl; l.ldc = E; l.R = xl; l.ldr = E; l.Cmap = &h->mc_xl;
                  l.M = B; l.N = E; l.K = E; FFB_TRY(launch_tc(h, l, stop, s)); }
                cur = xl; rows = B; Pq = 1;
            }
            FFB_TRY(launch_ln_split(h, cur, Lw.n2w, Lw.n2b, nullptr, ax2p, ssE, qpos_cross, qmod_cross, rows, E, stop, s));
            { TcLin l; l.A0 = &TS.m_x2p; l.W = &Tw.ca_q; l.w_scale = Tw.s_ca_q; l.bias = Lw.ca.in_b; l.M = rows; l.N = E; l.K = E;
              if (hp) { l.Cs = aqc; l.cs_stride = h->cap_rows * E; l.ldcs = E; l.Cmap = &h->ms_qc; }
              else { l.C = qkv; l.ldc = E; l.Cmap = &h->mc_qkv1; }
              FFB_TRY(launch_tc(h, l, stop, s)); }
            { AttnGroups g{}; g.ragged = 1; g.q_begin = seq_off; g.q_mul = Pq; g.k_begin = row_off; g.k_len = vlen;
              if (hp && h->attn_x_ok && (h->opt_attn_x & 1)) {
                  ax::Params ap{}; ap.mode = 0; ap.seq_off = seq_off; ap.q_mul = Pq; ap.row_off = row_off; ap.vlen = vlen;
                  ap.n_groups = N; ap.q_col = 0; ap.k_col = li * E; ap.v_col = li * E;
                  ap.Os = aatt; ap.os_stride = ssE; ap.ldo = E;
                  FFB_TRY(launch_attn_x(h, h->mx_q, h->mx_k, h->mx_vrow, ap, &h->h_seq_off, h->sum_seq_vlen * Pq,
                                        PC_ATTN_TILED, stop, s));
              } else if (hp) {
                  const long long kvs = (long long)h->R * LdE;
                  AttnHalfIn in{aqc, h->cap_rows * E, E, h->kc_h.as<uint16_t>() + (size_t)li * E, kvs, h->vc_h.as<uint16_t>() + (size_t)li * E, kvs, LdE};
                  FFB_TRY(launch_attn_h(h, in, aatt, ssE, g, N, h->max_seq_per_wf * Pq, h->max_vlen, h->sum_seq_vlen * Pq, PC_ATTN_TILED, stop, s));
              } else
              FFB_TRY(launch_attn_tiled(h, qkv, E, h->Kc.as<float>() + (size_t)li * E, h->Vc.as<float>() + (size_t)li * E, LdE,
                                        att, E, g, N, h->max_seq_per_wf * Pq, h->sum_seq_vlen * Pq, stop, s, aatt, ssE)); }
            { TcLin l; l.A0 = &TS.m_att; l.W = &Tw.ca_out; l.w_scale = Tw.s_ca_out; l.bias = Lw.ca.out_b; l.C = cur; l.ldc = E; l.R = cur; l.ldr = E; l.Cmap = (cur == x) ? &h->mc_x : &h->mc_xl;
              l.M = rows; l.N = E; l.K = E; FFB_TRY(launch_tc(h, l, stop, s)); }
            FFB_TRY(launch_ln_split(h, cur, Lw.n3w, Lw.n3b, ax2, nullptr, ssE, nullptr, 1, rows, E, stop, s));
            { TcLin l; l.A0 = &TS.m_x2; l.W = &Tw.l1; l.w_scale = Tw.s_l1; l.bias = Lw.l1b; l.relu = 1; l.Cs = ah; l.cs_stride = ssF; l.ldcs = FF; l.Cmap = &h->ms_h;
              l.M = rows; l.N = FF; l.K = E; FFB_TRY(launch_tc(h, l, stop, s)); }
            { TcLin l; l.A0 = &TS.m_h; l.W = &Tw.l2; l.w_scale = Tw.s_l2; l.bias = Lw.l2b; l.C = cur; l.ldc = E; l.R = cur; l.ldr = E; l.Cmap = (cur == x) ? &h->mc_x : &h->mc_xl;
              l.M = rows; l.N = E; l.K = FF; FFB_TRY(launch_tc(h, l, stop, s)); }
        }
    }
    // decoder.norm (transformer.py:115-116) + project (model_para.py:225) + select_next (model_para.py:173-179)
    const bool head64 = h->opt_head64 != 0;
    if (head64)        // last position of every sequence: LayerNorm in float64 (rounded to fp32); project is folded into memW (run_head_fold)
        FFB_TRY(launch_ln64(h, nullptr, cur, Pq, Pq - 1, w.dec_nw, w.dec_nb, nullptr, nullptr, h->hy32.as<float>(), nullptr, nullptr, 1, B, stop, s));
    if (head64 && h->opt_prune) {
    } else if (!tc) {
        FFB_TRY(launch_ln(h, cur, w.dec_nw, w.dec_nb, x2, rows, E, stop, s));
        { Lin l; l.A = x2; l.lda = E; l.W = w.proj_w; l.ldw = E; l.bias = w.proj_b; l.C = att; l.ldc = E; l.M = rows; l.N = E; l.K = E;
          FFB_TRY(launch_linear(h, l, stop, s)); }
    } else {
        FFB_TRY(launch_ln_split(h, cur, w.dec_nw, w.dec_nb, ax2, nullptr, ssE, nullptr, 1, rows, E, stop, s));
        { TcLin l; l.A0 = &TS.m_x2; l.W = &TS.proj; l.w_scale = TS.s_proj; l.bias = w.proj_b; l.C = att; l.ldc = E; l.Cmap = &h->mc_att; l.M = rows; l.N = E; l.K = E;
          FFB_TRY(launch_tc(h, l, stop, s)); }
    }
    PointerArgs pa{};
    pa.mem = h->mem.as<float>(); pa.ptr = att; pa.ptr_stride_rows = Pq; pa.ptr_off = Pq - 1;
    pa.ldm = E; pa.bias_col = 0;
    if (head64) { pa.mem = h->memW.as<float>(); pa.ldm = E + 4; pa.bias_col = E; pa.ptr = h->hy32.as<float>(); pa.ptr_stride_rows = 1; pa.ptr_off = 0; }
    pa.row_off = row_off; pa.v_len = vlen; pa.seq_wf = h->d_seq_wf.as<int>();
    pa.logits = h->logits.as<float>(); pa.L = h->L;
    const bool beam = append && h->W > 1;
    pa.tok_out = (append && !beam) ? tok_cur(h) + (size_t)P * B : nullptr;
    pa.B = B; pa.E = E; pa.num_token = h->cfg.num_token;
    pa.nonstop_count = (h->cfg.mode == FFB_MODE_PARALLEL && !beam) ? st + 2 : nullptr;
    pa.eos_count = (h->cfg.mode == FFB_MODE_SEQ2SEQ) ? st + 3 : nullptr;
    pa.stop = stop;
    prof_begin(h, PC_POINTER, 2.0 * E * h->sum_seq_vlen, s);
    const dim3 pb_grid((h->max_seq_per_wf + PB_SEQ - 1) / PB_SEQ, N);
    if (h->opt_pointer_batched && h->E % 128 == 0 && h->E <= 1024) {  // the sequences of a wireframe share its memory rows: stream them once per PB_SEQ sequences
        if (head64) launch_k(h, pointer_batched_kernel<true>, pb_grid, dim3(256), (size_t)PB_SEQ * E * sizeof(float), s, pa, seq_off);
        else launch_k(h, pointer_batched_kernel<false>, pb_grid, dim3(256), (size_t)PB_SEQ * E * sizeof(float), s, pa, seq_off);
    } else if (head64) launch_k(h, pointer_kernel<true>, dim3(B), dim3(256), 0, s, pa);
    else launch_k(h, pointer_kernel<false>, dim3(B), dim3(256), 0, s, pa);
    prof_end(h, s);
    h->launches++; CU(h, cudaGetLastError());
    if (beam) {      // select the W best continuations per anchor, re-order the token histories into the other buffer
        int* tok_new = h->tok.as<int>() + (size_t)(h->tok_sel ^ 1) * h->T * B;
        launch_k(h, beam_step_kernel, dim3(B / h->W), dim3(32 * h->W), 0, s, (const float*)h->logits.as<float>(), h->L, (const int*)h->d_seq_wf.as<int>(),
                 vlen, h->beam_cum.as<double>(), (const int*)tok_cur(h), tok_new, P, h->T, B, h->W, h->cfg.num_token, st + 2, stop);
        h->launches++; CU(h, cudaGetLastError());
        h->tok_sel ^= 1;
    }
    if (append && h->xchg_on) {
        XchgArgs xa{};
        for (int r = 0; r < h->xchg_world; ++r) xa.peers[r] = h->xchg_peers[r];
        xa.rank = h->xchg_rank; xa.world = h->xchg_world; xa.T = h->T;
        launch_k(h, step_end_xchg_kernel, dim3(1), dim3(32), 0, s, st + 2, st, st + 1, xa, P - 1, h->xchg_epoch, st + 7);
        h->launches++; CU(h, cudaGetLastError());
    } else if (append) {
        launch_k(h, step_end_kernel, dim3(1), dim3(1), 0, s, h->cfg.mode, B, st + 2, st + 3, st, st + 1);
        h->launches++; CU(h, cudaGetLastError());
    }
    h->last_P = P;
    return FFB_OK;
}


// ---- the whole greedy loop in one persistent cooperative kernel (persist.cuh) ------------------------------------------------------
constexpr int FFB_PD_UNAVAILABLE = 1;         // run_persistent: the cooperative launch was refused; the caller runs the per-step kernels instead
constexpr long long PD_AUTO_ROWS = 896;      // auto mode: batches with at most this many decoder rows (sequences x (T - 1)) in the last step; measured crossover
                                             // with the multi-kernel path at 860-1010 rows (profiles/probe_persist_threshold_r2.json)

// Decode steps the persistent kernel takes for the encoded batch (0 = none).  Hybrid mode (3): the steps whose row count (sequences x prefix
// length) stays in the latency-bound regime; the rest of the loop runs on the per-step kernels (the two share tok / state, so the hand-over is
// free -- but the layer-0 q / k / v cache is lost for the rest of that decode).
int persist_steps(const ffb_handle* h) {
    if (!h->opt_persist || h->pd_grid <= 0) return 0;
    if (h->tc_fmt != 2 || !h->tc_ok || !h->opt_tc) return 0;
    const ffb_handle::TcSet& TS = h->tcs[0];
    if (!TS.ready || (int)TS.layers.size() != h->Ld) return 0;
    if (!h->opt_head64 || h->W != 1 || h->xchg_on || h->encode_only) return 0;
    if (h->N > pd::MAX_WF || h->Ld > pd::MAX_LAYERS || h->E % 128 != 0 || h->E > 1024 || h->FF % pd::TK != 0 || h->H * 64 != h->E) return 0;
    if (h->opt_persist == 2) return h->T - 1;
    // auto / hybrid: never when a test forces one of the multi-kernel pipelines
    if (h->opt_tc != 1 || h->B <= 0) return 0;
    if (h->opt_persist == 3) return (int)std::min<long long>(h->T - 1, PD_AUTO_ROWS / h->B);     // hybrid: the first steps only
    return (h->B * (long long)(h->T - 1) <= PD_AUTO_ROWS) ? h->T - 1 : 0;                         // auto: whole decodes that stay below the crossover
}

int run_persistent(ffb_handle* h, int max_steps, cudaStream_t s) {
    const ffb_handle::TcSet& TS = h->tcs[0];
    pd::Params p{};
    for (int li = 0; li < h->Ld; ++li) {
        const DecLayerW& Lw = h->w.dec[li];
        const ffb_handle::DecTcW& Tw = TS.layers[li];
        pd::LayerP& L = p.L[li];
        L.w_sa_in = Tw.p_sa_in; L.w_sa_out = Tw.p_sa_out; L.w_ca_q = Tw.p_ca_q; L.w_ca_out = Tw.p_ca_out; L.w_l1 = Tw.p_l1; L.w_l2 = Tw.p_l2;
        L.s_sa_in = 1.0f / Tw.s_sa_in; L.s_sa_out = 1.0f / Tw.s_sa_out; L.s_ca_q = 1.0f / Tw.s_ca_q; L.s_ca_out = 1.0f / Tw.s_ca_out;
        L.s_l1 = 1.0f / Tw.s_l1; L.s_l2 = 1.0f / Tw.s_l2;
        L.b_sa_in = Lw.sa.in_b; L.b_sa_out = Lw.sa.out_b; L.b_ca_q = Lw.ca.in_b; L.b_ca_out = Lw.ca.out_b; L.b_l1 = Lw.l1b; L.b_l2 = Lw.l2b;
        L.n1w = Lw.n1w; L.n1b = Lw.n1b; L.n2w = Lw.n2w; L.n2b = Lw.n2b; L.n3w = Lw.n3w; L.n3b = Lw.n3b;
    }
    p.Ld = h->Ld; p.E = h->E; p.FF = h->FF; p.H = h->H; p.B = (int)h->B; p.T = h->T; p.N = h->N; p.Lrows = h->L; p.mode = h->cfg.mode;
    p.num_token = h->cfg.num_token; p.max_steps = max_steps;
    p.x = h->x.as<float>(); p.qkv = h->qkv.as<float>();
    p.xs = h->a_x2.as<uint16_t>(); p.xps = h->a_x2p.as<uint16_t>(); p.atts = h->a_att.as<uint16_t>(); p.hs = h->a_h.as<uint16_t>();
    p.ssE = h->cap_rows * h->E; p.ssF = h->cap_rows * h->FF;
    p.mem = h->mem.as<float>(); p.memW = h->memW.as<float>(); p.Kc = h->Kc.as<float>(); p.Vc = h->Vc.as<float>();
    p.qpos = h->w.qpos; p.dec_nw = h->w.dec_nw; p.dec_nb = h->w.dec_nb;
    p.row_off = h->d_row_off.as<int>(); p.vlen = h->d_vlen.as<int>(); p.seq_wf = h->d_seq_wf.as<int>(); p.seq_off = h->d_seq_off.as<int>();
    p.tok = tok_cur(h); p.logits = h->logits.as<float>(); p.state = h->state.as<int>();
    const size_t sync_bytes = (size_t)(h->T + 1) * sizeof(int);
    CU(h, h->pd_sync.ensure(sync_bytes));
    CU(h, cudaMemsetAsync(h->pd_sync.p, 0, sync_bytes, s));
    p.bar = h->pd_sync.as<unsigned>(); p.counts = h->pd_sync.as<int>() + 1;
    const bool prof = getenv("FFB_PD_PROF") != nullptr;          // debugging aid: per-phase clock sums of CTA 0 on stderr
    if (prof) {
        CU(h, h->pd_prof.ensure(64 * sizeof(long long)));
        CU(h, cudaMemsetAsync(h->pd_prof.p, 0, 64 * sizeof(long long), s));
        p.prof = h->pd_prof.as<long long>();
    }
    void* args[] = {(void*)&p};
    prof_begin(h, PC_OTHER, 0.0, s);
    const void* kern = (h->E <= 512) ? (const void*)pd::decode_persistent_kernel<4> : (const void*)pd::decode_persistent_kernel<8>;
    const cudaError_t le = cudaLaunchCooperativeKernel(kern, dim3(h->pd_grid), dim3(pd::THREADS), args, (size_t)pd::SMEM_BYTES, s);
    prof_end(h, s);
    if (le == cudaErrorCooperativeLaunchTooLarge || le == cudaErrorLaunchOutOfResources) {
        // the grid cannot be co-resident right now (another context holds SM resources): this handle stays on the multi-kernel path
        cudaGetLastError();
        h->pd_grid = 0;
        return FFB_PD_UNAVAILABLE;
    }
    CU(h, le);
    h->launches++;
    h->last_P = 0;
    if (prof) {
        long long v[64];
        CU(h, cudaMemcpyAsync(v, h->pd_prof.p, sizeof v, cudaMemcpyDeviceToHost, s));
        CU(h, cudaStreamSynchronize(s));
        static const char* names[11] = {"ln1", "qkv", "self_attn", "sa_out", "ln2", "cross", "ca_out", "ln3", "ffn1", "ffn2", "head"};
        for (int i = 0; i < 11; ++i) fprintf(stderr, "[pd] %-9s work %10lld clk   barrier %10lld clk\n", names[i], v[i], v[16 + i]);
        if (v[31] > 0) fprintf(stderr, "[pd] in-situ load latency (avg clk): activation via L2 %lld, parameter (read-only path) %lld, second activation %lld\n", v[24] / v[31], v[25] / v[31], v[23] / v[31]);
        if (v[31] > 0)
            fprintf(stderr, "[pd] LayerNorm row of CTA 0 warp 0 (avg clk over %lld rows): loads %lld, reductions %lld, format + store %lld\n", v[31], v[28] / v[31],
                    v[29] / v[31], v[30] / v[31]);
        for (int o = 32; o <= 40; o += 8)
            if (v[o + 5] > 0)
                fprintf(stderr, "[pd] %s attention item, CTA 0 warp 0 (avg clk over %lld items): Q + first tile landed %lld, S + softmax %lld, P V (+ later tiles) %lld, partials + barrier %lld, merge + store %lld\n",
                        o == 32 ? "self" : "cross", v[o + 5], v[o] / v[o + 5], v[o + 1] / v[o + 5], v[o + 2] / v[o + 5], v[o + 3] / v[o + 5], v[o + 4] / v[o + 5]);
        if (v[52] > 0)
            fprintf(stderr, "[pd] head of CTA 0 (avg clk over %lld sequences): row -> smem %lld, float64 LayerNorm %lld, memory-row scan %lld, merge + append %lld\n", v[52],
                    v[48] / v[52], v[49] / v[52], v[50] / v[52], v[51] / v[52]);
        if (v[27] > 0)
            fprintf(stderr, "[pd] residual-projection item of CTA 0 (avg clk over %lld items): issue %lld, first pair landed %lld, mainloop %lld, hand-over %lld, epilogue %lld\n",
                    v[27], v[11] / v[27], v[12] / v[27], v[13] / v[27], v[14] / v[27], v[15] / v[27]);
    }
    return FFB_OK;
}

// seq2seq 'pointer' output after a persistent decode: att <- project(decoder.norm(x)) for every row the loop can have produced (row-wise
// operations: rows beyond the executed steps hold stale but finite values and are never read back)
int run_project_rows(ffb_handle* h, cudaStream_t s) {
    const int E = h->E, rows = (int)(h->B * (long long)(h->T - 1));
    const ffb_handle::TcSet& TS = h->tcs[h->tc_fmt - 2];
    FFB_TRY(launch_ln_split(h, h->x.as<float>(), h->w.dec_nw, h->w.dec_nb, h->a_x2.as<uint16_t>(), nullptr, h->cap_rows * E, nullptr, 1, rows, E, nullptr, s));
    TcLin l; l.A0 = &TS.m_x2; l.W = &TS.proj; l.w_scale = TS.s_proj; l.bias = h->w.proj_b; l.C = h->att.as<float>(); l.ldc = E; l.Cmap = &h->mc_att;
    l.M = rows; l.N = E; l.K = E;
    return launch_tc(h, l, nullptr, s);
}

int copy_out(ffb_handle* h, const void* dev_src, void* dst, size_t bytes, int loc, cudaStream_t s) {
    CU(h, cudaMemcpyAsync(dst, dev_src, bytes, loc == FFB_HOST ? cudaMemcpyDeviceToHost : cudaMemcpyDeviceToDevice, s));
    if (loc == FFB_HOST) CU(h, cudaStreamSynchronize(s));
    return FFB_OK;
}

}  // namespace

// ======================================================================================================
extern "C" {

size_t ffb_weight_count(const ffb_config* cfg) { return validate_config(cfg) ? 0 : weight_count(cfg); }

const char* ffb_last_error(const ffb_handle* h) { return h ? h->err.c_str() : g_create_error.c_str(); }

int ffb_create(const ffb_config* cfg, ffb_handle** out) {
    if (!out) return fail(nullptr, FFB_ERR_ARG, "out is NULL");
    *out = nullptr;
    if (const char* why = validate_config(cfg)) return fail(nullptr, FFB_ERR_ARG, "invalid config: %s", why);
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev <= 0)
        return fail(nullptr, FFB_ERR_CUDA, "no CUDA device available (%s): libffb200 has no CPU fallback", cudaGetErrorString(e));
    if (cfg->device >= ndev) return fail(nullptr, FFB_ERR_ARG, "device %d out of range (%d devices)", cfg->device, ndev);
    e = cudaSetDevice(cfg->device);
    if (e != cudaSuccess) return fail(nullptr, FFB_ERR_CUDA, "cudaSetDevice(%d): %s", cfg->device, cudaGetErrorString(e));
    cudaDeviceProp prop;
    e = cudaGetDeviceProperties(&prop, cfg->device);
    if (e != cudaSuccess) return fail(nullptr, FFB_ERR_CUDA, "cudaGetDeviceProperties: %s", cudaGetErrorString(e));
    if (prop.major != 10)
        return fail(nullptr, FFB_ERR_UNSUPPORTED, "device %d is sm_%d%d; libffb200 is built for sm_100a only", cfg->device, prop.major, prop.minor);
    e = cudaFuncSetAttribute(attn_tiled_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, AT_SMEM_BYTES);
    if (e != cudaSuccess) return fail(nullptr, FFB_ERR_CUDA, "cudaFuncSetAttribute(attn_tiled): %s", cudaGetErrorString(e));
    e = cudaFuncSetAttribute(attn_h_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, AF_SMEM_BYTES);
    if (e != cudaSuccess) return fail(nullptr, FFB_ERR_CUDA, "cudaFuncSetAttribute(attn_h_kernel): %s", cudaGetErrorString(e));
    e = cudaFuncSetAttribute(attn_f16_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, AF_SMEM_BYTES);
    if (e != cudaSuccess) return fail(nullptr, FFB_ERR_CUDA, "cudaFuncSetAttribute(attn_f16_kernel): %s", cudaGetErrorString(e));
    e = cudaFuncSetAttribute(attn_mma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, AM_SMEM_BYTES);
    if (e != cudaSuccess) return fail(nullptr, FFB_ERR_CUDA, "cudaFuncSetAttribute(attn_mma_kernel): %s", cudaGetErrorString(e));
    e = cudaFuncSetAttribute(tc::gemm_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, tc::Cfg<2>::SMEM_BYTES);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(tc::gemm_kernel<2, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, tc::Cfg<2, 1>::SMEM_BYTES);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(tc::gemm_kernel<2, 1, 64>, cudaFuncAttributeMaxDynamicSharedMemorySize, tc::Cfg<2, 1, 64>::SMEM_BYTES);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(al::attn_long_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, al::SMEM_BYTES);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(tc2::gemm2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, tc2::SMEM_BYTES);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(tc::gemm_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, tc::Cfg<3>::SMEM_BYTES);
    if (e != cudaSuccess) return fail(nullptr, FFB_ERR_CUDA, "cudaFuncSetAttribute(tc::gemm_kernel): %s", cudaGetErrorString(e));
    e = cudaFuncSetAttribute(pointer_batched_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, PB_SEQ * 1024 * (int)sizeof(float));
    if (e == cudaSuccess) e = cudaFuncSetAttribute(pointer_batched_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, PB_SEQ * 1024 * (int)sizeof(float));
    if (e != cudaSuccess) return fail(nullptr, FFB_ERR_CUDA, "cudaFuncSetAttribute(pointer_batched_kernel): %s", cudaGetErrorString(e));
    e = cudaFuncSetAttribute(ax::attn_x_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, ax::SMEM_BYTES);
    if (e != cudaSuccess) return fail(nullptr, FFB_ERR_CUDA, "cudaFuncSetAttribute(attn_x_kernel): %s", cudaGetErrorString(e));
    if (!g_encode_tiled) {
        void* fn = nullptr;
        cudaDriverEntryPointQueryResult qres;
        e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
        if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess || !fn)
            return fail(nullptr, FFB_ERR_CUDA, "cudaGetDriverEntryPoint(cuTensorMapEncodeTiled) failed: %s", cudaGetErrorString(e));
        g_encode_tiled = (EncodeTiledFn)fn;
    }
    ffb_handle* h = new (std::nothrow) ffb_handle();
    if (!h) return fail(nullptr, FFB_ERR_ARG, "out of host memory");
    h->num_sms = prop.multiProcessorCount;
    h->tc_ok = (cfg->num_model % tc::BN == 0) && (cfg->num_feedforward % tc::BN == 0);
    h->cfg = *cfg;
    h->E = cfg->num_model; h->H = cfg->num_head; h->FF = cfg->num_feedforward;
    h->L = cfg->num_lines + cfg->num_token; h->T = cfg->seq_len;
    h->Le = cfg->num_encoder_layers; h->Ld = cfg->num_decoder_layers;
    h->opt_prune = (cfg->mode == FFB_MODE_PARALLEL) ? 1 : 0;
    {   // persistent decode kernel: the grid must be co-resident (grid-wide barriers), so it is sized from the occupancy of THIS device
        int coop = 0, per_sm = 0;
        cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, cfg->device);
        if (coop && cudaFuncSetAttribute(pd::decode_persistent_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, pd::SMEM_BYTES) == cudaSuccess &&
            cudaFuncSetAttribute(pd::decode_persistent_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, pd::SMEM_BYTES) == cudaSuccess &&
            cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, pd::decode_persistent_kernel<8>, pd::THREADS, (size_t)pd::SMEM_BYTES) == cudaSuccess)
            h->pd_grid = std::min(per_sm, 1) * h->num_sms;
        cudaGetLastError();
    }
    for (auto& ev : h->ev) cudaEventCreate(&ev);
    *out = h;
    return FFB_OK;
}

int ffb_destroy(ffb_handle* h) {
    if (!h) return FFB_OK;
    cudaSetDevice(h->cfg.device);
    cudaDeviceSynchronize();
    DevBuf* bufs[] = {&h->wblob, &h->wcross, &h->d_row_off, &h->d_vlen, &h->d_pos_idx, &h->d_edge_src, &h->d_edge_dst, &h->d_seq_wf,
                      &h->d_seq_first, &h->d_seq_off, &h->d_slot_seq, &h->d_seq_slot, &h->d_coords, &h->d_predict, &h->d_out_stage,
                      &h->d_mask_stage, &h->d_prefix, &h->mem, &h->Kc, &h->Vc, &h->tok, &h->logits, &h->state, &h->x, &h->x2, &h->qkv,
                      &h->att, &h->hb, &h->xl, &h->tcs[0].wsplit, &h->tcs[1].wsplit, &h->a_x2, &h->a_x2p, &h->a_att, &h->a_h, &h->a_qkv, &h->a_qc, &h->kc_h, &h->vc_h, &h->a_ql, &h->beam_cum, &h->x64, &h->y64, &h->yp64, &h->qkv64, &h->att64, &h->h64, &h->projT, &h->memW, &h->hy32, &h->d_tile_off, &h->e0pad, &h->a_c, &h->qkv0_cache, &h->a_qkv0, &h->pd_sync, &h->pd_prof, &h->d_label, &h->d_label_mask, &h->d_kmask, &h->dn_q_begin, &h->dn_edge_dst, &h->dn_pos_idx};
    for (DevBuf* b : bufs) b->release();
    for (auto& ev : h->ev) if (ev) cudaEventDestroy(ev);
    for (auto& ev : h->prof_pool) cudaEventDestroy(ev);
    if (h->h_stop) cudaFreeHost((void*)h->h_stop);
    ffb_stop_exchange_disconnect(h);
    if (h->xchg_buf) cudaFree(h->xchg_buf);
    delete h;
    return FFB_OK;
}

int ffb_set_option(ffb_handle* h, int option, int value) {
    if (!h) return FFB_ERR_ARG;
    switch (option) {
        case FFB_OPT_DEDUP_PAD: h->opt_dedup = value ? 1 : 0; h->encoded = false; return FFB_OK;
        case FFB_OPT_PRUNE_LAST: h->opt_prune = value ? 1 : 0; return FFB_OK;
        case FFB_OPT_ENCODER_PRECISION:
            if (value != 0 && value != 2) return fail(h, FFB_ERR_ARG, "FFB_OPT_ENCODER_PRECISION: 0 = fp16x2 tcgen05 / fp32, 2 = float64");
            h->opt_enc_prec = value; h->encoded = false; return FFB_OK;
        case FFB_OPT_HEAD_FP64: h->opt_head64 = value ? 1 : 0; return FFB_OK;
        case FFB_OPT_L0_CACHE: h->opt_l0cache = value ? 1 : 0; h->encoded = false; return FFB_OK;
        case FFB_OPT_PERSISTENT:
            if (value < 0 || value > 3) return fail(h, FFB_ERR_ARG, "FFB_OPT_PERSISTENT: 0 off, 1 auto, 2 wherever supported, 3 hybrid");
            h->opt_persist = value; return FFB_OK;
        case FFB_OPT_SKINNY_GEMM: h->opt_skinny = value ? 1 : 0; return FFB_OK;
        case FFB_OPT_ATTN_LONG: h->opt_attn_long = value ? 1 : 0; return FFB_OK;
        case FFB_OPT_ENCODE_ONLY: h->opt_encode_only = value ? 1 : 0; h->encoded = false; return FFB_OK;
        case FFB_OPT_FORCE_F:
            if (value < 0 || value > h->cfg.num_lines) return fail(h, FFB_ERR_ARG, "FFB_OPT_FORCE_F must be in [0, num_lines]");
            h->opt_force_F = value; h->encoded = false; return FFB_OK;
        case FFB_OPT_BEAM:
            if (value < 1 || value > BEAM_MAX) return fail(h, FFB_ERR_ARG, "FFB_OPT_BEAM: beam width must be in [1, %d]", BEAM_MAX);
            if (value > 1 && h->cfg.mode != FFB_MODE_PARALLEL) return fail(h, FFB_ERR_UNSUPPORTED, "beam search is specified for the parallel model only");
            h->opt_beam = value; h->encoded = false; return FFB_OK;
        case FFB_OPT_TIMING: h->opt_timing = value ? 1 : 0; return FFB_OK;
        case FFB_OPT_PROFILE: h->opt_profile = value ? 1 : 0; h->prof_recs.clear(); return FFB_OK;
        case FFB_OPT_TC_FORMAT:
            if (value != 2 && value != 3) return fail(h, FFB_ERR_ARG, "FFB_OPT_TC_FORMAT: 2 = fp16x2, 3 = bf16x3");
            h->tc_fmt = value; h->encoded = false;
            if (h->weights_loaded && h->tc_ok && h->opt_tc) { int rc = prepare_tc(h, value, nullptr); if (rc != FFB_OK) return rc; }
            return FFB_OK;
        case FFB_OPT_STAGGER: h->opt_stagger = value ? 1 : 0; return FFB_OK;
        case FFB_OPT_TMA_EPILOGUE: h->opt_tma_out = value ? 1 : 0; h->encoded = false; return FFB_OK;   // half_pipe is planned per batch
        case FFB_OPT_ATTN_MMA:
            if (value < 0 || value > 2) return fail(h, FFB_ERR_ARG, "FFB_OPT_ATTN_MMA: 0 SIMT, 1 3xTF32, 2 fp16x2");
            h->opt_attn_mma = value; h->encoded = false; return FFB_OK;
        case FFB_OPT_ATTN_X: h->opt_attn_x = value & 3; return FFB_OK;
        case FFB_OPT_ENCODER_TC: h->opt_enc_tc = value ? 1 : 0; return FFB_OK;
        case FFB_OPT_PDL:
            if (value < 0 || value > 2) return fail(h, FFB_ERR_ARG, "FFB_OPT_PDL: 0 off, 1 on, 2 auto");
            h->opt_pdl = value; h->pdl_on = (value == 1); return FFB_OK;
        case FFB_OPT_POINTER_BATCHED: h->opt_pointer_batched = value ? 1 : 0; return FFB_OK;
        case FFB_OPT_GEMM_VARIANT: if (value < 0 || value > 3) return fail(h, FFB_ERR_ARG, "FFB_OPT_GEMM_VARIANT: 0, 1, 2 (auto) or 3 (CTA pairs)"); h->opt_gemm_variant = value; return FFB_OK;
        case FFB_OPT_TENSOR_CORE:
            if (value < 0 || value > 2) return fail(h, FFB_ERR_ARG, "FFB_OPT_TENSOR_CORE: 0 off, 1 auto, 2 force");
            if (value && !h->tc_ok) return fail(h, FFB_ERR_UNSUPPORTED, "tensor-core path needs num_model and num_feedforward multiples of 256");
            h->opt_tc = value; h->encoded = false; return FFB_OK;
        default: return fail(h, FFB_ERR_ARG, "unknown option %d", option);
    }
}

int ffb_load_weights(ffb_handle* h, const float* blob, size_t count, int loc, void* stream) {
    if (!h) return FFB_ERR_ARG;
    if (!blob) return fail(h, FFB_ERR_ARG, "blob is NULL");
    const size_t want = weight_count(&h->cfg);
    if (count != want) return fail(h, FFB_ERR_ARG, "weight blob has %zu floats, config needs %zu (strict state_dict layout)", count, want);
    FFB_TRY(set_device(h));
    cudaStream_t s = (cudaStream_t)stream;
    CU(h, h->wblob.ensure(want * sizeof(float)));
    CU(h, cudaMemcpyAsync(h->wblob.p, blob, want * sizeof(float), loc == FFB_HOST ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToDevice, s));
    bind_weights(h);
    // pack cross-attention K / V projection weights of all decoder layers: [Ld*E, E] (+ biases [Ld*E])
    const size_t E = h->E, Ld = h->Ld;
    CU(h, h->wcross.ensure((2 * Ld * E * E + 2 * Ld * E) * sizeof(float)));
    float* ckw = h->wcross.as<float>(); float* cvw = ckw + Ld * E * E; float* ckb = cvw + Ld * E * E; float* cvb = ckb + Ld * E;
    for (size_t l = 0; l < Ld; ++l) {
        const AttnW& ca = h->w.dec[l].ca;       // in_proj rows: [0,E) q, [E,2E) k, [2E,3E) v (torch functional.py:5866-5873)
        CU(h, cudaMemcpyAsync(ckw + l * E * E, ca.in_w + E * E, E * E * sizeof(float), cudaMemcpyDeviceToDevice, s));
        CU(h, cudaMemcpyAsync(cvw + l * E * E, ca.in_w + 2 * E * E, E * E * sizeof(float), cudaMemcpyDeviceToDevice, s));
        CU(h, cudaMemcpyAsync(ckb + l * E, ca.in_b + E, E * sizeof(float), cudaMemcpyDeviceToDevice, s));
        CU(h, cudaMemcpyAsync(cvb + l * E, ca.in_b + 2 * E, E * sizeof(float), cudaMemcpyDeviceToDevice, s));
    }
    h->w.ckw = ckw; h->w.cvw = cvw; h->w.ckb = ckb; h->w.cvb = cvb;
    CU(h, h->projT.ensure((E + 1) * E * sizeof(float)));            // rows 0..E-1 = W_project^T, row E = b_project (folded head)
    transpose_kernel<<<grid1d((long long)(E * E)), 256, 0, s>>>(h->w.proj_w, h->projT.as<float>(), (int)E, (int)E);
    h->launches++; CU(h, cudaGetLastError());
    CU(h, cudaMemcpyAsync(h->projT.as<float>() + E * E, h->w.proj_b, E * sizeof(float), cudaMemcpyDeviceToDevice, s));
    if (h->cfg.in_dim <= 128) {                             // first embedding linear with K zero-padded to 128 (tensor-core encoder)
        CU(h, h->e0pad.ensure(E * 128 * sizeof(float)));
        pad_cols_kernel<<<grid1d((long long)(E * 128)), 256, 0, s>>>(h->w.e0w, h->e0pad.as<float>(), (int)E, h->cfg.in_dim, 128);
        h->launches++; CU(h, cudaGetLastError());
    }
    for (auto& T : h->tcs) T.ready = false;                 // split weights are rebuilt lazily per operand format
    if (h->tc_ok && h->opt_tc) FFB_TRY(prepare_tc(h, h->tc_fmt, s));
    CU(h, cudaStreamSynchronize(s));
    h->weights_loaded = true;
    h->encoded = false; h->decoded = false;                 // mem / Kc / Vc of an earlier ffb_encode belong to the old weights
    return FFB_OK;
}

int ffb_encode(ffb_handle* h, const float* coords, const uint8_t* pad_mask, const int64_t* num_input,
               int32_t N, int loc, void* stream) {
    if (!h) return FFB_ERR_ARG;
    h->encoded = false; h->decoded = false;
    if (!h->weights_loaded) return fail(h, FFB_ERR_STATE, "ffb_encode before ffb_load_weights");
    if (!coords || !pad_mask) return fail(h, FFB_ERR_ARG, "coords / pad_mask is NULL");
    if (N < 1) return fail(h, FFB_ERR_ARG, "n_wireframes must be >= 1");
    FFB_TRY(set_device(h));
    cudaStream_t s = (cudaStream_t)stream;
    const size_t nl = h->cfg.num_lines;
    const uint8_t* mask_h = pad_mask; const int64_t* ni_h = num_input;
    if (loc == FFB_DEVICE) {       // small control data is planned on the host (the reference syncs here too, model_para.py:187)
        h->h_mask.resize((size_t)N * nl);
        CU(h, cudaMemcpyAsync(h->h_mask.data(), pad_mask, (size_t)N * nl, cudaMemcpyDeviceToHost, s));
        if (num_input) {
            h->h_num_input.resize(N);
            CU(h, cudaMemcpyAsync(h->h_num_input.data(), num_input, (size_t)N * sizeof(int64_t), cudaMemcpyDeviceToHost, s));
            ni_h = h->h_num_input.data();
        }
        CU(h, cudaStreamSynchronize(s));
        mask_h = h->h_mask.data();
    }
    if (h->opt_timing) CU(h, cudaEventRecord(h->ev[0], s));
    FFB_TRY(plan_batch(h, mask_h, ni_h, N, s));
    const float* coords_dev = coords;
    if (loc == FFB_HOST) {
        const size_t bytes = (size_t)N * nl * h->cfg.in_dim * sizeof(float);
        CU(h, h->d_coords.ensure(bytes));
        CU(h, cudaMemcpyAsync(h->d_coords.p, coords, bytes, cudaMemcpyHostToDevice, s));
        coords_dev = h->d_coords.as<float>();
    }
    h->attn_allow_f16 = false;
    h->enc_used_64 = (h->opt_enc_prec == 2 && h->max_vlen <= ENC64_MAX_VLEN);
    int enc_rc = h->enc_used_64 ? run_encoder64(h, coords_dev, s) : run_encoder(h, coords_dev, s, true);
    if (h->enc_used_64) h->enc_used_tc = false;
    if (enc_rc == FFB_OK && h->enc_used_tc) {          // an activation left the fp16 range in the tensor-core encoder: redo it in fp32
        int ovf = 0;
        cudaError_t ce = cudaMemcpyAsync(&ovf, h->state.as<int>() + 6, sizeof(int), cudaMemcpyDeviceToHost, s);
        if (ce == cudaSuccess) ce = cudaStreamSynchronize(s);
        if (ce != cudaSuccess) { h->attn_allow_f16 = true; return fail(h, FFB_ERR_CUDA, "ffb_encode: %s", cudaGetErrorString(ce)); }
        if (ovf) { h->fp16_fallbacks++; enc_rc = run_encoder(h, coords_dev, s, false); }
    }
    h->attn_allow_f16 = true;
    FFB_TRY(enc_rc);
    if (h->opt_timing) CU(h, cudaEventRecord(h->ev[1], s));
    h->encoded = true;
    return FFB_OK;
}

int ffb_batch_info(const ffb_handle* h, int32_t* N, int32_t* F, int64_t* B, int64_t* B_eff, int64_t* R) {
    if (!h || !h->encoded) return FFB_ERR_STATE;
    if (N) *N = h->N;
    if (F) *F = h->F;
    if (B) *B = h->B_full;
    if (B_eff) *B_eff = h->B;
    if (R) *R = h->R;
    return FFB_OK;
}

int ffb_decode_greedy(ffb_handle* h, int64_t* predict, int loc, int32_t* steps_run, void* stream) {
    if (!h) return FFB_ERR_ARG;
    if (!h->encoded) return fail(h, FFB_ERR_STATE, "ffb_decode_greedy before ffb_encode");
    if (h->encode_only) return fail(h, FFB_ERR_STATE, "the batch was encoded with FFB_OPT_ENCODE_ONLY: no decode workspaces exist");
    if (!predict) return fail(h, FFB_ERR_ARG, "predict is NULL");
    FFB_TRY(set_device(h));
    cudaStream_t s = (cudaStream_t)stream;
    const int B = (int)h->B, T = h->T;
    int* st = h->state.as<int>();
    if (h->opt_timing && !h->ev[1]) return fail(h, FFB_ERR_STATE, "timing events missing");
    const long long n_slots = h->B_full;
    long long* out_dev = reinterpret_cast<long long*>(predict);
    if (loc == FFB_HOST) {
        CU(h, h->d_predict.ensure((size_t)n_slots * T * sizeof(long long)));
        out_dev = h->d_predict.as<long long>();
    }
    const bool syncing = (steps_run != nullptr) || (loc == FFB_HOST);
    if (h->xchg_on) {
        if (h->cfg.mode != FFB_MODE_PARALLEL) return fail(h, FFB_ERR_UNSUPPORTED, "batch splitting is implemented for the parallel model");
        h->xchg_epoch++;                           // every rank decodes the same number of times: epochs stay aligned
        CU(h, cudaMemsetAsync(st + 7, 0, sizeof(int), s));
    }
    for (int attempt = 0; attempt < 2; ++attempt) {
        CU(h, cudaMemsetAsync(st, 0, 5 * sizeof(int), s));          // [5] = overflow seen while encoding: survives
        h->tok_sel = 0;
        init_tokens_kernel<<<(B + 255) / 256, 256, 0, s>>>(h->d_seq_first.as<int>(), h->tok.as<int>(), B, st, st + 1, st + 3);
        h->launches++; CU(h, cudaGetLastError());
        if (h->W > 1) {
            beam_init_kernel<<<(B + 255) / 256, 256, 0, s>>>(h->beam_cum.as<double>(), B, h->W);
            h->launches++; CU(h, cudaGetLastError());
        }
        // No host sync inside the loop.  The stop flag is mirrored into pinned host memory after every step; once a copy that has
        // already landed shows it set, the remaining steps (whose kernels would all exit at once) are not launched at all.
        if (h->h_stop_cap < T) {
            if (h->h_stop) cudaFreeHost((void*)h->h_stop);
            h->h_stop = nullptr; h->h_stop_cap = 0;
            void* hp = nullptr;
            CU(h, cudaHostAlloc(&hp, (size_t)T * sizeof(int), cudaHostAllocDefault));
            h->h_stop = (volatile int*)hp; h->h_stop_cap = T;
        }
        for (int i = 0; i < T; ++i) h->h_stop[i] = 0;
        h->steps_launched = 0;
        int pd_steps = persist_steps(h);
        h->used_persist = pd_steps > 0;
        if (h->used_persist) {
            // small batch: the loop (or its first steps, while sequences x prefix length stays in the latency-bound regime) inside ONE
            // cooperative kernel; the stop predicate never leaves the device
            const int prc = run_persistent(h, pd_steps, s);
            if (prc == FFB_PD_UNAVAILABLE) { h->used_persist = false; pd_steps = 0; }
            else if (prc != FFB_OK) return prc;
        }
        if (h->used_persist) {
            if (!h->opt_prune) FFB_TRY(run_project_rows(h, s));       // seq2seq 'pointer' (model.py:216-217): project(decoder.norm(.)) of every position
            h->steps_launched = pd_steps;
            if (pd_steps < T - 1) CU(h, cudaMemcpyAsync((void*)(h->h_stop + pd_steps - 1), st, sizeof(int), cudaMemcpyDeviceToHost, s));
        }
        h->l0_suspend = h->used_persist;              // the layer-0 q / k / v cache holds nothing for the positions the persistent kernel decoded
        for (int step = pd_steps; step < T - 1; ++step) {
            bool stopped = false;
            for (int i = 0; i < step && !stopped; ++i) stopped = h->h_stop[i] != 0;
            if (stopped) break;
            FFB_TRY(run_step(h, step + 1, true, s));
            CU(h, cudaMemcpyAsync((void*)(h->h_stop + step), st, sizeof(int), cudaMemcpyDeviceToHost, s));
            h->steps_launched = step + 1;
        }
        expand_predict_kernel<<<grid1d(n_slots * T), 256, 0, s>>>(tok_cur(h), h->d_slot_seq.as<int>(), st + 1, out_dev, n_slots, B, T);
        h->launches++; CU(h, cudaGetLastError());
        if (!syncing) break;                      // fully asynchronous call: the caller must check ffb_overflowed() after its own sync
        int host_state[3] = {0, 0, 0};            // executed steps, fp16 overflow flag (decode), fp16 overflow flag (K/V cache split)
        CU(h, cudaMemcpyAsync(&host_state[0], st + 1, sizeof(int), cudaMemcpyDeviceToHost, s));
        CU(h, cudaMemcpyAsync(&host_state[1], st + 4, 2 * sizeof(int), cudaMemcpyDeviceToHost, s));
        CU(h, cudaStreamSynchronize(s));
        if (steps_run) *steps_run = host_state[0];
        if (h->xchg_on) {
            int xerr = 0;
            CU(h, cudaMemcpy(&xerr, st + 7, sizeof(int), cudaMemcpyDeviceToHost));
            if (xerr) return fail(h, FFB_ERR_CUDA, "stop-predicate exchange timed out: a peer rank did not reach decode step %d", host_state[0]);
        }
        if ((host_state[1] == 0 && !(h->half_pipe && host_state[2])) || h->tc_fmt != 2) break;
        if (h->xchg_on)      // a private re-run would desynchronise the ranks' step barriers
            return fail(h, FFB_ERR_UNSUPPORTED, "fp16 range overflow while decoding a split batch: set FFB_OPT_TC_FORMAT = 3 on every rank and run the batch again");
        // an activation left the fp16 range: switch this handle to the bf16x3 operand format (sticky) and decode again
        h->tc_fmt = 3; h->fp16_fallbacks++;
        FFB_TRY(prepare_tc(h, 3, s));
    }
    if (h->opt_timing) CU(h, cudaEventRecord(h->ev[2], s));
    if (loc == FFB_HOST) {
        CU(h, cudaMemcpyAsync(predict, out_dev, (size_t)n_slots * T * sizeof(long long), cudaMemcpyDeviceToHost, s));
        CU(h, cudaStreamSynchronize(s));
    }
    h->decoded = true;
    return FFB_OK;
}

int ffb_forward_eval(ffb_handle* h, const float* coords, const uint8_t* pad_mask, const int64_t* num_input, int32_t N,
                     int64_t* predict, int loc, int32_t* steps_run, void* stream) {
    FFB_TRY(ffb_encode(h, coords, pad_mask, num_input, N, loc, stream));
    return ffb_decode_greedy(h, predict, loc, steps_run, stream);
}

int ffb_featurize(ffb_handle* h, const double* points, const int64_t* edge_off, const int64_t* wf_edge_off, int32_t N,
                  float* coords, uint8_t* pad_mask, int64_t* num_input, int loc, void* stream) {
    if (!h) return FFB_ERR_ARG;
    if (!points || !edge_off || !wf_edge_off || !coords || !pad_mask || !num_input) return fail(h, FFB_ERR_ARG, "ffb_featurize: NULL argument");
    if (N < 1) return fail(h, FFB_ERR_ARG, "n_wireframes must be >= 1");
    if (h->cfg.in_dim % 2 != 0) return fail(h, FFB_ERR_UNSUPPORTED, "ffb_featurize needs point_dim 2");
    FFB_TRY(set_device(h));
    cudaStream_t s = (cudaStream_t)stream;
    const int nl = h->cfg.num_lines, P = h->cfg.in_dim / 2;
    if (P < 2) return fail(h, FFB_ERR_UNSUPPORTED, "ffb_featurize needs >= 2 points per line");
    // the offsets are planned on the host (they are tiny); validate like the reference would fail (IndexError / empty edge)
    std::vector<int64_t> wf(N + 1);
    CU(h, cudaMemcpyAsync(wf.data(), wf_edge_off, (N + 1) * sizeof(int64_t), loc == FFB_HOST ? cudaMemcpyHostToHost : cudaMemcpyDeviceToHost, s));
    CU(h, cudaStreamSynchronize(s));
    if (wf[0] != 0) return fail(h, FFB_ERR_ARG, "wf_edge_off[0] must be 0");
    for (int i = 0; i < N; ++i) {
        if (wf[i + 1] < wf[i]) return fail(h, FFB_ERR_ARG, "wf_edge_off must be non-decreasing");
        if (wf[i + 1] - wf[i] > nl) return fail(h, FFB_ERR_ARG, "wireframe %d has %lld edges, num_lines is %d (the reference raises IndexError)", i, (long long)(wf[i + 1] - wf[i]), nl);
    }
    const int64_t ne = wf[N];
    std::vector<int64_t> eo(ne + 1);
    CU(h, cudaMemcpyAsync(eo.data(), edge_off, (ne + 1) * sizeof(int64_t), loc == FFB_HOST ? cudaMemcpyHostToHost : cudaMemcpyDeviceToHost, s));
    CU(h, cudaStreamSynchronize(s));
    if (eo[0] != 0) return fail(h, FFB_ERR_ARG, "edge_off[0] must be 0");
    for (int64_t e = 0; e < ne; ++e)
        if (eo[e + 1] - eo[e] < 1) return fail(h, FFB_ERR_ARG, "edge %lld has no points", (long long)e);
    const int64_t npts = eo[ne];
    const size_t out_bytes = (size_t)N * nl * P * 2 * sizeof(float);
    const double* d_pts = points; const int64_t* d_eo = edge_off; const int64_t* d_wf = wf_edge_off;
    float* d_out = coords; uint8_t* d_mask = pad_mask; int64_t* d_ni = num_input;
    DevBuf in_stage, out_stage;
    if (loc == FFB_HOST) {
        const size_t pb = (size_t)std::max<int64_t>(npts, 1) * 2 * sizeof(double), eb = (size_t)(ne + 1) * sizeof(int64_t), wb = (size_t)(N + 1) * sizeof(int64_t);
        const size_t a8 = (pb + 7) / 8 * 8;
        if (in_stage.ensure(a8 + eb + wb) != cudaSuccess || out_stage.ensure(out_bytes + (size_t)N * sizeof(int64_t) + (size_t)N * nl) != cudaSuccess) {
            in_stage.release(); out_stage.release();
            return fail(h, FFB_ERR_CUDA, "ffb_featurize: out of device memory");
        }
        uint8_t* ib = in_stage.as<uint8_t>();
        cudaMemcpyAsync(ib, points, (size_t)npts * 2 * sizeof(double), cudaMemcpyHostToDevice, s);
        cudaMemcpyAsync(ib + a8, eo.data(), eb, cudaMemcpyHostToDevice, s);
        cudaMemcpyAsync(ib + a8 + eb, wf.data(), wb, cudaMemcpyHostToDevice, s);
        d_pts = reinterpret_cast<const double*>(ib); d_eo = reinterpret_cast<const int64_t*>(ib + a8); d_wf = reinterpret_cast<const int64_t*>(ib + a8 + eb);
        uint8_t* ob = out_stage.as<uint8_t>();
        d_out = reinterpret_cast<float*>(ob); d_ni = reinterpret_cast<int64_t*>(ob + out_bytes); d_mask = ob + out_bytes + (size_t)N * sizeof(int64_t);
    }
    featurize_kernel<<<grid1d((long long)N * nl * P), 256, 0, s>>>(d_pts, reinterpret_cast<const long long*>(d_eo), reinterpret_cast<const long long*>(d_wf),
                                                                d_out, d_mask, reinterpret_cast<long long*>(d_ni), N, nl, P);
    h->launches++;
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess && loc == FFB_HOST) {
        cudaMemcpyAsync(coords, d_out, out_bytes, cudaMemcpyDeviceToHost, s);
        cudaMemcpyAsync(num_input, d_ni, (size_t)N * sizeof(int64_t), cudaMemcpyDeviceToHost, s);
        cudaMemcpyAsync(pad_mask, d_mask, (size_t)N * nl, cudaMemcpyDeviceToHost, s);
        e = cudaStreamSynchronize(s);
    }
    if (loc == FFB_HOST) { if (e == cudaSuccess) e = cudaStreamSynchronize(s); in_stage.release(); out_stage.release(); }
    if (e != cudaSuccess) return fail(h, FFB_ERR_CUDA, "ffb_featurize: %s", cudaGetErrorString(e));
    return FFB_OK;
}

int ffb_parse_faces(ffb_handle* h, const int64_t* predict, int32_t N, int32_t F, const double* points, const int64_t* edge_off,
                    const int64_t* wf_edge_off, double tol, int32_t check_enclosed, uint8_t* valid, int32_t* face_type, int32_t* n_loops,
                    int32_t* loop_len, int32_t* indices, int32_t* n_indices, int loc, void* stream) {
    if (!h) return FFB_ERR_ARG;
    if (!predict || !points || !edge_off || !wf_edge_off || !valid || !face_type || !n_loops || !loop_len || !indices || !n_indices)
        return fail(h, FFB_ERR_ARG, "ffb_parse_faces: NULL argument");
    if (N < 1 || F < 1) return fail(h, FFB_ERR_ARG, "ffb_parse_faces: n_wireframes and F must be >= 1");
    const int T = h->T;
    if (T > PF_MAX_T) return fail(h, FFB_ERR_UNSUPPORTED, "ffb_parse_faces: seq_len %d > %d", T, PF_MAX_T);
    FFB_TRY(set_device(h));
    cudaStream_t s = (cudaStream_t)stream;
    const size_t S = (size_t)N * F;
    const long long* d_pred = reinterpret_cast<const long long*>(predict);
    const double* d_pts = points; const int64_t* d_eo = edge_off; const int64_t* d_wf = wf_edge_off;
    uint8_t* d_valid = valid; int* d_ft = face_type; int* d_nl = n_loops; int* d_ll = loop_len; int* d_idx = indices; int* d_ni = n_indices;
    DevBuf in_stage, out_stage;
    size_t o_ft = 0, o_nl = 0, o_ni = 0, o_ll = 0, o_idx = 0, o_valid = 0;
    if (loc == FFB_HOST) {
        std::vector<int64_t> wf(wf_edge_off, wf_edge_off + N + 1);
        const int64_t ne = wf[N];
        const int64_t npts = edge_off[ne];
        const size_t pb = (size_t)std::max<int64_t>(npts, 1) * 2 * sizeof(double), eb = (size_t)(ne + 1) * 8, wb = (size_t)(N + 1) * 8, prb = S * T * 8;
        o_ft = 0; o_nl = o_ft + S * 4; o_ni = o_nl + S * 4; o_ll = o_ni + S * 4; o_idx = o_ll + S * T * 4; o_valid = o_idx + S * T * 4;
        if (in_stage.ensure(pb + eb + wb + prb) != cudaSuccess || out_stage.ensure(o_valid + S) != cudaSuccess) {
            in_stage.release(); out_stage.release();
            return fail(h, FFB_ERR_CUDA, "ffb_parse_faces: out of device memory");
        }
        uint8_t* ib = in_stage.as<uint8_t>();
        cudaMemcpyAsync(ib, points, (size_t)npts * 16, cudaMemcpyHostToDevice, s);
        cudaMemcpyAsync(ib + pb, edge_off, eb, cudaMemcpyHostToDevice, s);
        cudaMemcpyAsync(ib + pb + eb, wf_edge_off, wb, cudaMemcpyHostToDevice, s);
        cudaMemcpyAsync(ib + pb + eb + wb, predict, prb, cudaMemcpyHostToDevice, s);
        d_pts = reinterpret_cast<const double*>(ib); d_eo = reinterpret_cast<const int64_t*>(ib + pb); d_wf = reinterpret_cast<const int64_t*>(ib + pb + eb);
        d_pred = reinterpret_cast<const long long*>(ib + pb + eb + wb);
        uint8_t* ob = out_stage.as<uint8_t>();
        d_ft = reinterpret_cast<int*>(ob + o_ft); d_nl = reinterpret_cast<int*>(ob + o_nl); d_ni = reinterpret_cast<int*>(ob + o_ni);
        d_ll = reinterpret_cast<int*>(ob + o_ll); d_idx = reinterpret_cast<int*>(ob + o_idx); d_valid = ob + o_valid;
    }
    parse_faces_kernel<<<(unsigned)((S + 127) / 128), 128, 0, s>>>(d_pred, N, F, T, d_pts, reinterpret_cast<const long long*>(d_eo),
                                                                  reinterpret_cast<const long long*>(d_wf), tol, check_enclosed, h->cfg.num_token, 1,
                                                                  d_valid, d_ft, d_nl, d_ll, d_idx, d_ni);
    h->launches++;
    cudaError_t e = cudaGetLastError();
    if (loc == FFB_HOST) {
        if (e == cudaSuccess) {
            cudaMemcpyAsync(face_type, d_ft, S * 4, cudaMemcpyDeviceToHost, s); cudaMemcpyAsync(n_loops, d_nl, S * 4, cudaMemcpyDeviceToHost, s);
            cudaMemcpyAsync(n_indices, d_ni, S * 4, cudaMemcpyDeviceToHost, s); cudaMemcpyAsync(loop_len, d_ll, S * T * 4, cudaMemcpyDeviceToHost, s);
            cudaMemcpyAsync(indices, d_idx, S * T * 4, cudaMemcpyDeviceToHost, s); cudaMemcpyAsync(valid, d_valid, S, cudaMemcpyDeviceToHost, s);
        }
        const cudaError_t e2 = cudaStreamSynchronize(s);
        if (e == cudaSuccess) e = e2;
        in_stage.release(); out_stage.release();
    }
    if (e != cudaSuccess) return fail(h, FFB_ERR_CUDA, "ffb_parse_faces: %s", cudaGetErrorString(e));
    return FFB_OK;
}

int ffb_get_memory(ffb_handle* h, float* memory, int loc, void* stream) {
    if (!h) return FFB_ERR_ARG;
    if (!h->encoded) return fail(h, FFB_ERR_STATE, "ffb_get_memory before ffb_encode");
    if (!memory) return fail(h, FFB_ERR_ARG, "memory is NULL");
    FFB_TRY(set_device(h));
    cudaStream_t s = (cudaStream_t)stream;
    const size_t bytes = (size_t)h->N * h->L * h->E * sizeof(float);
    float* dst = memory;
    if (loc == FFB_HOST) { CU(h, h->d_out_stage.ensure(bytes)); dst = h->d_out_stage.as<float>(); }
    unpack_memory_kernel<<<grid1d((long long)h->N * h->L * (h->E / 4)), 256, 0, s>>>(h->mem.as<float>(), h->d_row_off.as<int>(),
                                                                                  h->d_vlen.as<int>(), dst, h->N, h->L, h->E);
    h->launches++; CU(h, cudaGetLastError());
    if (loc == FFB_HOST) FFB_TRY(copy_out(h, dst, memory, bytes, loc, s));
    return FFB_OK;
}

int ffb_set_memory(ffb_handle* h, const float* memory, int loc, void* stream) {
    if (!h) return FFB_ERR_ARG;
    if (!h->encoded) return fail(h, FFB_ERR_STATE, "ffb_set_memory before ffb_encode");
    if (!memory) return fail(h, FFB_ERR_ARG, "memory is NULL");
    FFB_TRY(set_device(h));
    cudaStream_t s = (cudaStream_t)stream;
    const size_t bytes = (size_t)h->N * h->L * h->E * sizeof(float);
    const float* src = memory;
    if (loc == FFB_HOST) {
        CU(h, h->d_out_stage.ensure(bytes));
        CU(h, cudaMemcpyAsync(h->d_out_stage.p, memory, bytes, cudaMemcpyHostToDevice, s));
        src = h->d_out_stage.as<float>();
    }
    pack_memory_kernel<<<grid1d((long long)h->R * (h->E / 4)), 256, 0, s>>>(src, h->d_row_off.as<int>(), h->d_vlen.as<int>(), h->mem.as<float>(),
                                                                         h->N, h->L, h->E);
    h->launches++; CU(h, cudaGetLastError());
    FFB_TRY(run_cross_cache(h, s, false));
    FFB_TRY(run_head_fold(h, s, false));
    FFB_TRY(run_cross_cache_split(h, s));
    h->decoded = false;
    return FFB_OK;
}

static int emit_logits(ffb_handle* h, float* logits, int loc, cudaStream_t s) {
    const size_t bytes = (size_t)h->B_full * h->L * sizeof(float);
    float* dst = logits;
    if (loc == FFB_HOST) { CU(h, h->d_out_stage.ensure(bytes)); dst = h->d_out_stage.as<float>(); }
    expand_rows_kernel<<<grid1d(h->B_full * h->L), 256, 0, s>>>(h->logits.as<float>(), h->d_slot_seq.as<int>(), dst, h->B_full, h->L);
    h->launches++; CU(h, cudaGetLastError());
    if (loc == FFB_HOST) FFB_TRY(copy_out(h, dst, logits, bytes, loc, s));
    return FFB_OK;
}

int ffb_get_last_logits(ffb_handle* h, float* logits, int loc, void* stream) {
    if (!h) return FFB_ERR_ARG;
    if (!h->decoded) return fail(h, FFB_ERR_STATE, "ffb_get_last_logits before a decode");
    if (!logits) return fail(h, FFB_ERR_ARG, "logits is NULL");
    FFB_TRY(set_device(h));
    return emit_logits(h, logits, loc, (cudaStream_t)stream);
}

int ffb_get_last_pointer(ffb_handle* h, float* pointer, int32_t* P_out, int loc, void* stream) {
    if (!h) return FFB_ERR_ARG;
    if (!h->decoded) return fail(h, FFB_ERR_STATE, "ffb_get_last_pointer before a decode");
    if (h->opt_prune) return fail(h, FFB_ERR_STATE, "ffb_get_last_pointer needs FFB_OPT_PRUNE_LAST = 0");
    if (!pointer) return fail(h, FFB_ERR_ARG, "pointer is NULL");
    FFB_TRY(set_device(h));
    cudaStream_t s = (cudaStream_t)stream;
    int steps = 0;
    CU(h, cudaMemcpyAsync(&steps, h->state.as<int>() + 1, sizeof(int), cudaMemcpyDeviceToHost, s));
    CU(h, cudaStreamSynchronize(s));
    if (steps < 1) return fail(h, FFB_ERR_STATE, "no decode step was executed");
    if (P_out) *P_out = steps;
    // rows are (sequence, position): exactly [N, P, E] for the last executed step (later steps exited early)
    return copy_out(h, h->att.p, pointer, (size_t)h->B * steps * h->E * sizeof(float), loc, s);
}

int ffb_forced_prefix_logits(ffb_handle* h, const int64_t* prefix, int32_t P, float* logits, int loc, void* stream) {
    if (!h) return FFB_ERR_ARG;
    if (!h->encoded) return fail(h, FFB_ERR_STATE, "ffb_forced_prefix_logits before ffb_encode");
    if (h->encode_only) return fail(h, FFB_ERR_STATE, "the batch was encoded with FFB_OPT_ENCODE_ONLY");
    if (!prefix || !logits) return fail(h, FFB_ERR_ARG, "prefix / logits is NULL");
    if (P < 1 || P > h->T - 1) return fail(h, FFB_ERR_ARG, "P must be in [1, T-1]");
    if (h->W > 1) return fail(h, FFB_ERR_STATE, "forced prefixes need FFB_OPT_BEAM = 1");
    if (h->B != h->B_full) return fail(h, FFB_ERR_STATE, "forced prefixes need FFB_OPT_DEDUP_PAD = 0 before ffb_encode");
    h->tok_sel = 0;
    FFB_TRY(set_device(h));
    cudaStream_t s = (cudaStream_t)stream;
    const size_t pbytes = (size_t)P * h->B_full * sizeof(int64_t);
    const long long* pdev = reinterpret_cast<const long long*>(prefix);
    if (loc == FFB_HOST) {
        CU(h, h->d_prefix.ensure(pbytes));
        CU(h, cudaMemcpyAsync(h->d_prefix.p, prefix, pbytes, cudaMemcpyHostToDevice, s));
        pdev = h->d_prefix.as<long long>();
    }
    // validate on the host that every token addresses an un-masked memory row of its wireframe
    {
        std::vector<int64_t> hp((size_t)P * h->B_full);
        CU(h, cudaMemcpyAsync(hp.data(), pdev, pbytes, cudaMemcpyDeviceToHost, s));
        CU(h, cudaStreamSynchronize(s));
        for (int p = 0; p < P; ++p)
            for (long long b = 0; b < h->B_full; ++b) {
                const int wf = (int)(b / h->F);
                const int64_t t = hp[(size_t)p * h->B_full + b];
                if (t < 0 || t >= h->h_vlen[wf]) return fail(h, FFB_ERR_ARG, "prefix[%d,%lld]=%lld is not an un-masked row", p, b, (long long)t);
            }
    }
    int* st = h->state.as<int>();
    CU(h, cudaMemsetAsync(st, 0, 5 * sizeof(int), s));
    load_prefix_kernel<<<grid1d((long long)P * h->B), 256, 0, s>>>(pdev, h->d_seq_slot.as<int>(), h->tok.as<int>(), P, (int)h->B_full, (int)h->B);
    h->launches++; CU(h, cudaGetLastError());
    for (int attempt = 0; attempt < 2; ++attempt) {
        FFB_TRY(run_step(h, P, false, s));
        int ovf[2] = {0, 0};
        CU(h, cudaMemcpyAsync(ovf, st + 4, 2 * sizeof(int), cudaMemcpyDeviceToHost, s));
        CU(h, cudaStreamSynchronize(s));
        if ((ovf[0] == 0 && !(h->half_pipe && ovf[1])) || h->tc_fmt != 2) break;
        h->tc_fmt = 3; h->fp16_fallbacks++;
        FFB_TRY(prepare_tc(h, 3, s));
        CU(h, cudaMemsetAsync(st, 0, 5 * sizeof(int), s));
    }
    h->decoded = false;
    return emit_logits(h, logits, loc, s);
}

// ---- forward_train's `embedding` output: the encoder memory of ALL L rows of every wireframe, padded edges included ---------------------
// The decode path packs the un-masked rows only (padded rows are never read there).  Trainer.compute_loss (trainer.py:61-66) however takes
// the softmax over ALL L rows of outputs['embedding'], so the rows of padded edges -- queries that attend to the un-masked keys like every
// other row, transformer.py:169-171 -- carry probability mass and must be the reference's values.  This is a second, dense fp32 encoder pass
// (rows i * L + r; keys = the un-masked prefix of the wireframe) on the SIMT kernels; the encoder is ~1 % of a teacher-forced pass.
int run_encoder_dense(ffb_handle* h, const float* coords_dev, float* out_dev, cudaStream_t s) {
    const int E = h->E, FF = h->FF, N = h->N, L = h->L, nl = h->cfg.num_lines, nt = h->cfg.num_token;
    const int R = N * L, Re = N * nl;
    const Weights& w = h->w;
    std::vector<int> q_begin(N + 1), edge_dst(Re), pos_idx(R);
    for (int i = 0; i <= N; ++i) q_begin[i] = i * L;
    for (int i = 0; i < N; ++i) {
        for (int e = 0; e < nl; ++e) edge_dst[(size_t)i * nl + e] = i * L + nt + e;
        for (int r = 0; r < L; ++r) pos_idx[(size_t)i * L + r] = r;
    }
    FFB_TRY(upload(h, h->dn_q_begin, q_begin, s));
    FFB_TRY(upload(h, h->dn_edge_dst, edge_dst, s));
    FFB_TRY(upload(h, h->dn_pos_idx, pos_idx, s));
    CU(h, cudaStreamSynchronize(s));                       // the std::vectors are pageable
    const size_t f4 = sizeof(float), rows = (size_t)std::max(R, Re);
    CU(h, h->x.ensure(rows * E * f4)); CU(h, h->x2.ensure(rows * E * f4)); CU(h, h->qkv.ensure(rows * 3 * E * f4));
    CU(h, h->att.ensure(rows * E * f4)); CU(h, h->hb.ensure(rows * std::max(FF, E) * f4));
    float* x = h->x.as<float>(); float* x2 = h->x2.as<float>(); float* qkv = h->qkv.as<float>(); float* att = h->att.as<float>(); float* hb = h->hb.as<float>();
    const int* qb = h->dn_q_begin.as<int>(); const int* pidx = h->dn_pos_idx.as<int>();
    // VanillaEmedding (embedding.py:23-38): every edge row, zero-padded ones included
    { Lin l; l.A = coords_dev; l.lda = h->cfg.in_dim; l.W = w.e0w; l.ldw = h->cfg.in_dim; l.bias = w.e0b; l.C = hb; l.ldc = E; l.M = Re; l.N = E; l.K = h->cfg.in_dim;
      l.relu = 1; FFB_TRY(launch_linear(h, l, nullptr, s)); }
    { Lin l; l.A = hb; l.lda = E; l.W = w.e2w; l.ldw = E; l.bias = w.e2b; l.C = x; l.ldc = E; l.c_rows = h->dn_edge_dst.as<int>(); l.M = Re; l.N = E; l.K = E;
      FFB_TRY(launch_linear(h, l, nullptr, s)); }
    token_rows_kernel<<<grid1d((long long)N * nt * (E / 4)), 256, 0, s>>>(w.tok_table, qb, x, N, nt, E);
    h->launches++; CU(h, cudaGetLastError());
    AttnGroups g{}; g.ragged = 1; g.q_begin = qb; g.q_mul = 1; g.k_begin = qb; g.k_len = h->d_vlen.as<int>();
    double qk = 0; for (int i = 0; i < N; ++i) qk += (double)L * h->h_vlen[i];
    for (int li = 0; li < h->Le; ++li) {                                  // TransformerEncoderLayer.forward_pre (transformer.py:164-176)
        const EncLayerW& Lw = w.enc[li];
        FFB_TRY(launch_ln(h, x, Lw.n1w, Lw.n1b, x2, R, E, nullptr, s));
        { Lin l; l.A = x2; l.lda = E; l.W = Lw.sa.in_w; l.ldw = E; l.bias = Lw.sa.in_b; l.C = qkv; l.ldc = 3 * E;
          l.pos = w.pos; l.ldpos = E; l.pos_idx = pidx; l.pos_cols = 2 * E; l.M = R; l.N = 3 * E; l.K = E; FFB_TRY(launch_linear(h, l, nullptr, s)); }
        FFB_TRY(launch_attn_tiled(h, qkv, 3 * E, qkv + E, qkv + 2 * E, 3 * E, att, E, g, N, L, qk, nullptr, s));
        { Lin l; l.A = att; l.lda = E; l.W = Lw.sa.out_w; l.ldw = E; l.bias = Lw.sa.out_b; l.C = x; l.ldc = E; l.R = x; l.ldr = E;
          l.M = R; l.N = E; l.K = E; FFB_TRY(launch_linear(h, l, nullptr, s)); }
        FFB_TRY(launch_ln(h, x, Lw.n2w, Lw.n2b, x2, R, E, nullptr, s));
        { Lin l; l.A = x2; l.lda = E; l.W = Lw.l1w; l.ldw = E; l.bias = Lw.l1b; l.C = hb; l.ldc = FF; l.M = R; l.N = FF; l.K = E; l.relu = 1;
          FFB_TRY(launch_linear(h, l, nullptr, s)); }
        { Lin l; l.A = hb; l.lda = FF; l.W = Lw.l2w; l.ldw = FF; l.bias = Lw.l2b; l.C = x; l.ldc = E; l.R = x; l.ldr = E;
          l.M = R; l.N = E; l.K = FF; FFB_TRY(launch_linear(h, l, nullptr, s)); }
    }
    return launch_ln(h, x, w.enc_nw, w.enc_nb, out_dev, R, E, nullptr, s);   // encoder.norm (transformer.py:80-81)
}

int ffb_forward_train(ffb_handle* h, const float* coords, const uint8_t* pad_mask, const int64_t* num_input, int32_t N, const int64_t* label,
                      const uint8_t* label_mask, int32_t label_rows, float* pointer, float* embedding, int32_t loc, void* stream) {
    if (!h) return FFB_ERR_ARG;
    if (!label || !label_mask || !pointer) return fail(h, FFB_ERR_ARG, "label / label_mask / pointer is NULL");
    if (h->xchg_on || h->opt_beam != 1) return fail(h, FFB_ERR_STATE, "ffb_forward_train needs beam width 1 and no batch splitting");
    // every (wireframe, anchor slot) owns its label sequence: no de-duplication of padded anchors, no last-layer pruning (all positions are outputs)
    const int dedup0 = h->opt_dedup, prune0 = h->opt_prune, force0 = h->opt_force_F;
    h->opt_dedup = 0; h->opt_prune = 0;
    int rc = ffb_encode(h, coords, pad_mask, num_input, N, loc, stream);
    h->opt_dedup = dedup0; h->opt_force_F = force0;
    if (rc != FFB_OK) { h->opt_prune = prune0; return rc; }
    cudaStream_t s = (cudaStream_t)stream;
    auto done = [&](int code) { h->opt_prune = prune0; h->train_kmask = nullptr; h->decoded = false; return code; };
    if (label_rows < h->F) return done(fail(h, FFB_ERR_ARG, "label_rows (%d) < F = max(num_input) (%d)", label_rows, h->F));
    const int T = h->T, P = T - 1, B = (int)h->B;
    const size_t n_lab = (size_t)N * label_rows * T;
    const long long* lab_dev = reinterpret_cast<const long long*>(label);
    const unsigned char* lm_dev = label_mask;
    std::vector<int64_t> hl(n_lab);
    if (loc == FFB_HOST) {
        if (h->d_label.ensure(n_lab * sizeof(int64_t)) != cudaSuccess || h->d_label_mask.ensure(n_lab) != cudaSuccess) return done(fail(h, FFB_ERR_CUDA, "out of device memory (labels)"));
        if (cudaMemcpyAsync(h->d_label.p, label, n_lab * sizeof(int64_t), cudaMemcpyHostToDevice, s) != cudaSuccess ||
            cudaMemcpyAsync(h->d_label_mask.p, label_mask, n_lab, cudaMemcpyHostToDevice, s) != cudaSuccess) return done(fail(h, FFB_ERR_CUDA, "label upload failed"));
        lab_dev = h->d_label.as<long long>(); lm_dev = h->d_label_mask.as<unsigned char>();
        memcpy(hl.data(), label, n_lab * sizeof(int64_t));
    } else {
        if (cudaMemcpyAsync(hl.data(), label, n_lab * sizeof(int64_t), cudaMemcpyDeviceToHost, s) != cudaSuccess || cudaStreamSynchronize(s) != cudaSuccess)
            return done(fail(h, FFB_ERR_CUDA, "label download failed"));
    }
    // every teacher token must address an un-masked memory row of its wireframe (the reference's torch.gather would read padding otherwise)
    for (int wf = 0; wf < N; ++wf)
        for (int f = 0; f < h->F; ++f)
            for (int p = 0; p < P; ++p) {
                const int64_t t = hl[((size_t)wf * label_rows + f) * T + p];
                if (t < 0 || t >= h->h_vlen[wf]) return done(fail(h, FFB_ERR_ARG, "label[%d,%d,%d]=%lld is not an un-masked memory row", wf, f, p, (long long)t));
            }
    if (h->d_kmask.ensure((size_t)B * P) != cudaSuccess) return done(fail(h, FFB_ERR_CUDA, "out of device memory (key mask)"));
    int* st = h->state.as<int>();
    if (cudaMemsetAsync(st, 0, 5 * sizeof(int), s) != cudaSuccess) return done(fail(h, FFB_ERR_CUDA, "memset failed"));
    h->tok_sel = 0;
    for (int attempt = 0; attempt < 2; ++attempt) {
        load_labels_kernel<<<grid1d((long long)B * P), 256, 0, s>>>(lab_dev, lm_dev, label_rows, T, h->d_seq_wf.as<int>(), h->d_seq_off.as<int>(), h->tok.as<int>(),
                                                                    h->d_kmask.as<unsigned char>(), B, P);
        h->launches++;
        h->train_kmask = h->d_kmask.as<unsigned char>();
        rc = run_step(h, P, false, s);
        h->train_kmask = nullptr;
        if (rc != FFB_OK) return done(rc);
        int ovf[2] = {0, 0};
        if (cudaMemcpyAsync(ovf, st + 4, 2 * sizeof(int), cudaMemcpyDeviceToHost, s) != cudaSuccess || cudaStreamSynchronize(s) != cudaSuccess)
            return done(fail(h, FFB_ERR_CUDA, "state download failed"));
        if ((ovf[0] == 0 && !(h->half_pipe && ovf[1])) || h->tc_fmt != 2) break;
        h->tc_fmt = 3; h->fp16_fallbacks++;                                   // an activation left the fp16 range: bf16x3 operands, once more
        rc = prepare_tc(h, 3, s);
        if (rc != FFB_OK) return done(rc);
        cudaMemsetAsync(st, 0, 5 * sizeof(int), s);
    }
    // project(decoder(...)) of every position: rows ordered (sequence, position) = [N * F, T - 1, E] (model_para.py:161,166)
    rc = copy_out(h, h->att.p, pointer, (size_t)B * P * h->E * sizeof(float), loc, s);
    if (rc == FFB_OK && embedding != nullptr) {
        // outputs['embedding'] before its replication per anchor slot: [N, L, E] with the rows of padded edges as the reference computes them
        const size_t eb = (size_t)N * h->L * h->E * sizeof(float);
        float* dst = embedding;
        if (loc == FFB_HOST) { if (h->d_out_stage.ensure(eb) != cudaSuccess) return done(fail(h, FFB_ERR_CUDA, "out of device memory (embedding)")); dst = h->d_out_stage.as<float>(); }
        const float* coords_dev = (loc == FFB_HOST) ? h->d_coords.as<float>() : coords;
        rc = run_encoder_dense(h, coords_dev, dst, s);
        if (rc == FFB_OK && loc == FFB_HOST) rc = copy_out(h, dst, embedding, eb, loc, s);
    }
    return done(rc);
}

int ffb_stop_exchange_export(ffb_handle* h, void* ipc_handle_out) {
    if (!h || !ipc_handle_out) return FFB_ERR_ARG;
    FFB_TRY(set_device(h));
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    if (!h->xchg_buf) {
        const size_t bytes = 2 * (size_t)h->T * XCHG_MAXW * sizeof(int);
        CU(h, cudaMalloc(&h->xchg_buf, bytes));
        CU(h, cudaMemset(h->xchg_buf, 0, bytes));
    }
    cudaIpcMemHandle_t hd;
    CU(h, cudaIpcGetMemHandle(&hd, h->xchg_buf));
    memcpy(ipc_handle_out, &hd, sizeof hd);
    return FFB_OK;
}

int ffb_stop_exchange_connect(ffb_handle* h, int32_t rank, int32_t world, const void* ipc_handles) {
    if (!h || !ipc_handles) return FFB_ERR_ARG;
    if (world < 1 || world > XCHG_MAXW || rank < 0 || rank >= world) return fail(h, FFB_ERR_ARG, "stop exchange: need 0 <= rank < world <= %d", XCHG_MAXW);
    if (!h->xchg_buf) return fail(h, FFB_ERR_STATE, "ffb_stop_exchange_connect before ffb_stop_exchange_export");
    if (h->xchg_on) return fail(h, FFB_ERR_STATE, "stop exchange is already connected");
    FFB_TRY(set_device(h));
    for (int r = 0; r < world; ++r) {
        if (r == rank) { h->xchg_peers[r] = h->xchg_buf; continue; }
        cudaIpcMemHandle_t hd;
        memcpy(&hd, reinterpret_cast<const uint8_t*>(ipc_handles) + (size_t)r * sizeof hd, sizeof hd);
        void* p = nullptr;
        cudaError_t e = cudaIpcOpenMemHandle(&p, hd, cudaIpcMemLazyEnablePeerAccess);
        if (e != cudaSuccess) {
            for (int q = 0; q < r; ++q) if (q != rank && h->xchg_peers[q]) { cudaIpcCloseMemHandle(h->xchg_peers[q]); h->xchg_peers[q] = nullptr; }
            return fail(h, FFB_ERR_CUDA, "cudaIpcOpenMemHandle(rank %d): %s", r, cudaGetErrorString(e));
        }
        h->xchg_peers[r] = static_cast<int*>(p);
    }
    h->xchg_rank = rank; h->xchg_world = world; h->xchg_epoch = 0; h->xchg_on = true;
    return FFB_OK;
}

int ffb_stop_exchange_disconnect(ffb_handle* h) {
    if (!h) return FFB_ERR_ARG;
    if (h->xchg_on) {
        cudaSetDevice(h->cfg.device);
        cudaDeviceSynchronize();
        for (int r = 0; r < h->xchg_world; ++r)
            if (r != h->xchg_rank && h->xchg_peers[r]) cudaIpcCloseMemHandle(h->xchg_peers[r]);
        for (auto& p : h->xchg_peers) p = nullptr;
        h->xchg_on = false;
    }
    return FFB_OK;
}

int ffb_get_beams(ffb_handle* h, int64_t* beams, double* scores, int loc, void* stream) {
    if (!h) return FFB_ERR_ARG;
    if (!h->decoded) return fail(h, FFB_ERR_STATE, "ffb_get_beams before a decode");
    if (!beams || !scores) return fail(h, FFB_ERR_ARG, "beams / scores is NULL");
    if (h->W < 2) return fail(h, FFB_ERR_STATE, "ffb_get_beams needs FFB_OPT_BEAM > 1");
    FFB_TRY(set_device(h));
    cudaStream_t s = (cudaStream_t)stream;
    const long long n_slots = h->B_full;
    const size_t nb = (size_t)n_slots * h->W * h->T * sizeof(long long), ns = (size_t)n_slots * h->W * sizeof(double);
    long long* bd = reinterpret_cast<long long*>(beams); double* sd = scores;
    if (loc == FFB_HOST) {
        CU(h, h->d_out_stage.ensure(nb + ns));
        bd = h->d_out_stage.as<long long>(); sd = reinterpret_cast<double*>(h->d_out_stage.as<uint8_t>() + nb);
    }
    expand_beams_kernel<<<grid1d(n_slots * h->W * h->T), 256, 0, s>>>(tok_cur(h), h->beam_cum.as<double>(), h->d_slot_seq.as<int>(), h->state.as<int>() + 1,
                                                                     bd, sd, n_slots, (int)h->B, h->W, h->T);
    h->launches++; CU(h, cudaGetLastError());
    if (loc == FFB_HOST) {
        CU(h, cudaMemcpyAsync(beams, bd, nb, cudaMemcpyDeviceToHost, s));
        CU(h, cudaMemcpyAsync(scores, sd, ns, cudaMemcpyDeviceToHost, s));
        CU(h, cudaStreamSynchronize(s));
    }
    return FFB_OK;
}

int ffb_overflowed(ffb_handle* h, int32_t* overflowed, void* stream) {
    if (!h || !overflowed) return FFB_ERR_ARG;
    *overflowed = 0;
    if (!h->decoded) return fail(h, FFB_ERR_STATE, "ffb_overflowed before a decode");
    FFB_TRY(set_device(h));
    cudaStream_t s = (cudaStream_t)stream;
    int ovf[2] = {0, 0};
    CU(h, cudaMemcpyAsync(ovf, h->state.as<int>() + 4, 2 * sizeof(int), cudaMemcpyDeviceToHost, s));
    CU(h, cudaStreamSynchronize(s));
    if (h->tc_fmt == 2 && (ovf[0] || (h->half_pipe && ovf[1]))) {
        *overflowed = 1;
        h->tc_fmt = 3; h->fp16_fallbacks++;            // sticky: every later encode / decode on this handle runs in bf16x3
        h->encoded = false; h->decoded = false;       // the predictions just produced are invalid: the caller must run the batch again
        FFB_TRY(prepare_tc(h, 3, s));
    }
    return FFB_OK;
}

int ffb_steps_launched(const ffb_handle* h) { return h ? h->steps_launched : 0; }

int ffb_used_persistent(const ffb_handle* h) { return (h && h->used_persist) ? 1 : 0; }

int64_t ffb_kernel_launches(const ffb_handle* h) { return h ? h->launches : 0; }

int ffb_fp16_fallbacks(const ffb_handle* h) { return h ? h->fp16_fallbacks : 0; }

int ffb_phase_times(ffb_handle* h, float* out_ms, int32_t n) {
    if (!h || !out_ms || n < 2) return FFB_ERR_ARG;
    if (!h->opt_timing || !h->decoded) return fail(h, FFB_ERR_STATE, "timing not enabled or no forward recorded");
    CU(h, cudaEventSynchronize(h->ev[2]));
    CU(h, cudaEventElapsedTime(&out_ms[0], h->ev[0], h->ev[1]));
    CU(h, cudaEventElapsedTime(&out_ms[1], h->ev[1], h->ev[2]));
    return FFB_OK;
}

int ffb_profile_read(ffb_handle* h, int32_t n_classes, float* ms, double* flops, int64_t* launches) {
    if (!h || !ms || !flops || !launches || n_classes < PC_COUNT) return FFB_ERR_ARG;
    FFB_TRY(set_device(h));
    CU(h, cudaDeviceSynchronize());
    for (int c = 0; c < n_classes; ++c) { ms[c] = 0.f; flops[c] = 0.0; launches[c] = 0; }
    for (size_t i = 0; i < h->prof_recs.size(); ++i) {
        float t = 0.f;
        CU(h, cudaEventElapsedTime(&t, h->prof_pool[2 * i], h->prof_pool[2 * i + 1]));
        const int c = h->prof_recs[i].cls;
        ms[c] += t; flops[c] += h->prof_recs[i].flops; launches[c] += 1;
    }
    h->prof_recs.clear();
    return FFB_OK;
}

// ---- op-level hooks ---------------------------------------------------------------------------------
int ffb_op_linear(ffb_handle* h, const float* A, const float* W, const float* bias, const float* R, const float* pos,
                  int32_t pos_mod, int32_t pos_cols, float* C, int32_t M, int32_t N, int32_t K, int32_t relu, void* stream) {
    if (!h) return FFB_ERR_ARG;
    FFB_TRY(set_device(h));
    Lin l; l.A = A; l.lda = K; l.W = W; l.ldw = K; l.bias = bias; l.C = C; l.ldc = N; l.R = R; l.ldr = N;
    l.pos = pos; l.ldpos = K; l.pos_mod = pos_mod; l.pos_cols = pos_cols; l.M = M; l.N = N; l.K = K; l.relu = relu;
    if (pos && pos_mod < 1) return fail(h, FFB_ERR_ARG, "op_linear: pos needs pos_mod >= 1");
    return launch_linear(h, l, nullptr, (cudaStream_t)stream);
}

int ffb_op_linear_tc(ffb_handle* h, const float* A, const float* W, const float* bias, const float* R, float* C,
                     int32_t M, int32_t N, int32_t K, int32_t relu, int32_t via_split, void* stream) {
    if (!h) return FFB_ERR_ARG;
    if (M < 1 || N % tc::BN != 0 || K % tc::BK != 0) return fail(h, FFB_ERR_ARG, "op_linear_tc: need M >= 1, N %% 256 == 0, K %% 32 == 0");
    FFB_TRY(set_device(h));
    cudaStream_t s = (cudaStream_t)stream;
    const size_t Mp = ((size_t)M + 127) / 128 * 128;
    DevBuf as, ws, cs;
    int rc = FFB_OK;
    do {
        if (as.ensure(3 * Mp * K * 2) != cudaSuccess || ws.ensure(3 * (size_t)N * K * 2) != cudaSuccess ||
            cs.ensure(3 * Mp * (size_t)N * 2) != cudaSuccess) { rc = fail(h, FFB_ERR_CUDA, "op_linear_tc: out of device memory"); break; }
        if (cudaMemsetAsync(as.p, 0, 3 * Mp * K * 2, s) != cudaSuccess) { rc = fail(h, FFB_ERR_CUDA, "memset failed"); break; }
        const int fmt = h->tc_fmt;
        {   // A is [M,K] contiguous: split it as one array, the splits are then M*K elements apart
            const long long n4 = (long long)M * K / 4;
            split_array_kernel<<<grid1d(n4), 256, 0, s>>>(A, as.as<uint16_t>(), n4, 1.0f, fmt);
            h->launches++;
        }
        CUtensorMap mA, mW;
        if ((rc = encode_operand_map(h, &mA, as.p, K, M, tc::BM, fmt)) != FFB_OK) break;
        float wscale = 1.f;
        if ((rc = split_weight(h, W, ws.as<uint16_t>(), N, K, &mW, fmt, &wscale, s)) != FFB_OK) break;
        TcLin l; l.A0 = &mA; l.W = &mW; l.w_scale = wscale; l.bias = bias; l.M = M; l.N = N; l.K = K; l.relu = relu;
        if (via_split) { l.Cs = cs.as<uint16_t>(); l.cs_stride = (long long)Mp * N; l.ldcs = N; }
        else { l.C = C; l.ldc = N; l.R = R; l.ldr = N; }
        CUtensorMap mC;
        if (!via_split) { if ((rc = encode_output_map(h, &mC, C, N, M)) != FFB_OK) break; l.Cmap = &mC; }
        else if (fmt == 2) { if ((rc = encode_split_store_map(h, &mC, cs.p, N, Mp)) != FFB_OK) break; l.Cmap = &mC; }
        if ((rc = launch_tc(h, l, nullptr, s)) != FFB_OK) break;
        if (via_split) {
            sum_split_kernel<<<grid1d((long long)M * N), 256, 0, s>>>(cs.as<uint16_t>(), (long long)Mp * N, C, (long long)M * N, fmt);
            h->launches++;
        }
        if (cudaStreamSynchronize(s) != cudaSuccess) { rc = fail(h, FFB_ERR_CUDA, "op_linear_tc: %s", cudaGetErrorString(cudaGetLastError())); break; }
    } while (0);
    as.release(); ws.release(); cs.release();
    return rc;
}

int ffb_bench_linear_tc(ffb_handle* h, int32_t M, int32_t N, int32_t K, int32_t flags, int32_t iters, float* ms_out, void* stream) {
    if (!h || !ms_out || iters < 1) return FFB_ERR_ARG;
    if (M < 1 || N % tc::BN != 0 || K % tc::BK != 0) return fail(h, FFB_ERR_ARG, "bench_linear_tc: need N %% 256 == 0, K %% 32 == 0");
    FFB_TRY(set_device(h));
    cudaStream_t s = (cudaStream_t)stream;
    const size_t Mp = ((size_t)M + 127) / 128 * 128;
    DevBuf as, ws, cs, cf, bias;
    int rc = FFB_OK;
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    do {
        if (as.ensure(3 * Mp * K * 2) != cudaSuccess || ws.ensure(3 * (size_t)N * K * 2) != cudaSuccess || bias.ensure((size_t)N * 4) != cudaSuccess ||
            cs.ensure(3 * Mp * (size_t)N * 2) != cudaSuccess || cf.ensure(Mp * (size_t)N * 4) != cudaSuccess) {
            rc = fail(h, FFB_ERR_CUDA, "bench_linear_tc: out of device memory"); break; }
        cudaMemsetAsync(as.p, 0, 3 * Mp * K * 2, s); cudaMemsetAsync(ws.p, 0, 3 * (size_t)N * K * 2, s);
        if (flags & 32) {   // realistic operand bits (power draw depends on the data): pseudo-random fp16 values in (-2, 2) and their small "lo" parts
            fill_random_half_kernel<<<grid1d((long long)(3 * Mp * K)), 256, 0, s>>>(as.as<uint16_t>(), (long long)(3 * Mp * K), (long long)(Mp * K), 1u);
            fill_random_half_kernel<<<grid1d((long long)(3 * (size_t)N * K)), 256, 0, s>>>(ws.as<uint16_t>(), (long long)(3 * (size_t)N * K), (long long)((size_t)N * K), 2u);
        }
        cudaMemsetAsync(cf.p, 0, Mp * (size_t)N * 4, s); cudaMemsetAsync(bias.p, 0, (size_t)N * 4, s);
        CUtensorMap mA, mW;
        if ((rc = encode_operand_map(h, &mA, as.p, K, Mp, tc::BM, h->tc_fmt)) != FFB_OK) break;
        if ((rc = encode_weight_maps(h, &mW, ws.p, K, N, h->tc_fmt)) != FFB_OK) break;
        TcLin l; l.A0 = &mA; l.W = &mW; l.M = M; l.N = N; l.K = K;
        if (flags & 1) l.bias = bias.as<float>();
        if (flags & 4) l.relu = 1;
        if (flags & 64) l.dry_store = 1;
        if (flags & 8) { l.Cs = cs.as<uint16_t>(); l.cs_stride = (long long)Mp * N; l.ldcs = N; }
        else if (!(flags & 16)) { l.C = cf.as<float>(); l.ldc = N; if (flags & 2) { l.R = cf.as<float>(); l.ldr = N; } }
        CUtensorMap mC;
        if (l.C) { if ((rc = encode_output_map(h, &mC, cf.p, N, Mp)) != FFB_OK) break; l.Cmap = &mC; }
        else if (l.Cs && h->tc_fmt == 2) { if ((rc = encode_split_store_map(h, &mC, cs.p, N, Mp)) != FFB_OK) break; l.Cmap = &mC; }
        cudaEventCreate(&e0); cudaEventCreate(&e1);
        if ((rc = launch_tc(h, l, nullptr, s)) != FFB_OK) break;          // warm-up
        cudaEventRecord(e0, s);
        for (int i = 0; i < iters && rc == FFB_OK; ++i) rc = launch_tc(h, l, nullptr, s);
        cudaEventRecord(e1, s);
        if (cudaEventSynchronize(e1) != cudaSuccess) { rc = fail(h, FFB_ERR_CUDA, "bench_linear_tc: %s", cudaGetErrorString(cudaGetLastError())); break; }
        float ms = 0.f; cudaEventElapsedTime(&ms, e0, e1);
        *ms_out = ms / iters;
    } while (0);
    if (e0) cudaEventDestroy(e0);
    if (e1) cudaEventDestroy(e1);
    as.release(); ws.release(); cs.release(); cf.release(); bias.release();
    return rc;
}

int ffb_op_layernorm(ffb_handle* h, const float* x, const float* gamma, const float* beta, float* y, int32_t M, int32_t E, void* stream) {
    if (!h) return FFB_ERR_ARG;
    if (E % 128 != 0 || E > 1024) return fail(h, FFB_ERR_ARG, "op_layernorm: E must be a multiple of 128, <= 1024");
    FFB_TRY(set_device(h));
    return launch_ln(h, x, gamma, beta, y, M, E, nullptr, (cudaStream_t)stream);
}

int ffb_op_attention(ffb_handle* h, int32_t kind, const float* q, int32_t ldq, const float* k, const float* v, int32_t ldk,
                     float* out, int32_t G, int32_t nq, int32_t nk, int32_t H, void* stream) {
    if (!h) return FFB_ERR_ARG;
    if (H != h->H) return fail(h, FFB_ERR_ARG, "op_attention: H must equal the handle's num_head");
    FFB_TRY(set_device(h));
    cudaStream_t s = (cudaStream_t)stream;
    if (kind == 4) {     // inputs converted to fp16x2 first, then the half-input kernel (k and v must share ldk; out is fp32)
        const size_t nqe = (size_t)G * nq * ldq, nke = (size_t)G * nk * ldk;
        if (ldq % 4 || ldk % 4) return fail(h, FFB_ERR_ARG, "op_attention kind 4: ldq/ldk must be multiples of 4");
        DevBuf qh, kh, vh;
        int rc = FFB_OK;
        if (qh.ensure(2 * nqe * 2) != cudaSuccess || kh.ensure(2 * nke * 2) != cudaSuccess || vh.ensure(2 * nke * 2) != cudaSuccess)
            rc = fail(h, FFB_ERR_CUDA, "op_attention: out of device memory");
        if (rc == FFB_OK) {
            split_array_kernel<<<grid1d((long long)nqe / 4), 256, 0, s>>>(q, qh.as<uint16_t>(), (long long)nqe / 4, 1.0f, 2);
            split_array_kernel<<<grid1d((long long)nke / 4), 256, 0, s>>>(k, kh.as<uint16_t>(), (long long)nke / 4, 1.0f, 2);
            split_array_kernel<<<grid1d((long long)nke / 4), 256, 0, s>>>(v, vh.as<uint16_t>(), (long long)nke / 4, 1.0f, 2);
            AttnHalfIn in{qh.as<uint16_t>(), (long long)nqe, ldq, kh.as<uint16_t>(), (long long)nke, vh.as<uint16_t>(), (long long)nke, ldk};
            AttnGroups g{}; g.ragged = 0; g.nq = nq; g.nk = nk; g.q_stride = nq; g.q_off = 0; g.k_stride = nk; g.o_stride = nq;
            const int saved_fmt = h->tc_fmt; h->tc_fmt = 2;
            rc = launch_attn_h(h, in, nullptr, 0, g, G, nq, nk, (double)G * nq * nk, PC_ATTN_TILED, nullptr, s, out);
            h->tc_fmt = saved_fmt;
            if (cudaStreamSynchronize(s) != cudaSuccess && rc == FFB_OK) rc = fail(h, FFB_ERR_CUDA, "op_attention: %s", cudaGetErrorString(cudaGetLastError()));
        }
        qh.release(); kh.release(); vh.release();
        return rc;
    }
    if (kind == 7) {
        // tcgen05 streaming kernel (attn_l.cuh): G groups of nq == nk rows, queries = keys rows of the group (encoder self-attention)
        if (nq != nk) return fail(h, FFB_ERR_UNSUPPORTED, "op_attention kind 7: needs nq == nk");
        if (ldq != H * 64 || ldk % 32 || ldk < H * 64) return fail(h, FFB_ERR_ARG, "op_attention kind 7: ldq must be H*64, ldk a multiple of 32");
        const size_t Rk = (size_t)G * nk;
        const int tiles_g = (nk + 2 * al::BQ - 1) / (2 * al::BQ);          // pairs of 128-query tiles per group
        std::vector<int> tile_off(G + 1), row_off(G), vlen(G);
        for (int g = 0; g <= G; ++g) tile_off[g] = g * tiles_g;
        for (int g = 0; g < G; ++g) { row_off[g] = g * nk; vlen[g] = nk; }
        DevBuf qh, kh, vh, os, meta;
        int rc = FFB_OK;
        do {
            if (qh.ensure(2 * Rk * ldq * 2) != cudaSuccess || kh.ensure(2 * Rk * ldk * 2) != cudaSuccess || vh.ensure(2 * Rk * ldk * 2) != cudaSuccess ||
                os.ensure(2 * Rk * ldq * 2) != cudaSuccess || meta.ensure((3 * (size_t)G + 1) * sizeof(int)) != cudaSuccess) {
                rc = fail(h, FFB_ERR_CUDA, "op_attention: out of device memory"); break; }
            int* m = meta.as<int>();
            cudaMemcpyAsync(m, tile_off.data(), (G + 1) * sizeof(int), cudaMemcpyHostToDevice, s);
            cudaMemcpyAsync(m + G + 1, row_off.data(), G * sizeof(int), cudaMemcpyHostToDevice, s);
            cudaMemcpyAsync(m + 2 * G + 1, vlen.data(), G * sizeof(int), cudaMemcpyHostToDevice, s);
            if (cudaStreamSynchronize(s) != cudaSuccess) { rc = fail(h, FFB_ERR_CUDA, "op_attention: upload failed"); break; }
            split_array_kernel<<<grid1d((long long)Rk * ldq / 4), 256, 0, s>>>(q, qh.as<uint16_t>(), (long long)Rk * ldq / 4, 1.0f, 2);
            split_array_kernel<<<grid1d((long long)Rk * ldk / 4), 256, 0, s>>>(k, kh.as<uint16_t>(), (long long)Rk * ldk / 4, 1.0f, 2);
            split_array_kernel<<<grid1d((long long)Rk * ldk / 4), 256, 0, s>>>(v, vh.as<uint16_t>(), (long long)Rk * ldk / 4, 1.0f, 2);
            h->launches += 3;
            CUtensorMap mq, mk, mv;
            if ((rc = encode_rows_map(h, &mq, qh.p, ldq, Rk, 32, al::BQ, CU_TENSOR_MAP_SWIZZLE_64B)) != FFB_OK) break;
            if ((rc = encode_rows_map(h, &mk, kh.p, ldk, Rk, 32, al::KC, CU_TENSOR_MAP_SWIZZLE_64B)) != FFB_OK) break;
            if ((rc = encode_rows_map(h, &mv, vh.p, ldk, Rk, 64, al::KC, CU_TENSOR_MAP_SWIZZLE_128B)) != FFB_OK) break;
            al::Params lp{};
            lp.tile_off = m; lp.row_off = m + G + 1; lp.vlen = m + 2 * G + 1; lp.n_groups = G;
            lp.Os = os.as<uint16_t>(); lp.os_stride = (long long)Rk * ldq; lp.ldo = ldq;
            if ((rc = launch_attn_long(h, mq, mk, mv, lp, (long long)G * tiles_g, (double)G * nq * nk, PC_ATTN_TILED, s)) != FFB_OK) break;
            sum_split_kernel<<<grid1d((long long)Rk * ldq), 256, 0, s>>>(os.as<uint16_t>(), (long long)Rk * ldq, out, (long long)Rk * ldq, 2);
            h->launches++;
            if (cudaStreamSynchronize(s) != cudaSuccess) { rc = fail(h, FFB_ERR_CUDA, "op_attention kind 7: %s", cudaGetErrorString(cudaGetLastError())); break; }
        } while (0);
        qh.release(); kh.release(); vh.release(); os.release(); meta.release();
        return rc;
    }
    if (kind == 5 || kind == 6) {
        // tcgen05 kernel.  5: CROSS (group g = "wireframe" g with nq queries, nk <= 256 keys); 6: SELF (block-diagonal over G
        // sequences of nq == nk <= 128 rows)
        const bool self = (kind == 6);
        if (nk > ax::KMAX || (!self && G > ax::MAX_GROUPS)) return fail(h, FFB_ERR_UNSUPPORTED, "op_attention kind 5: needs nk <= 256 and G <= 255");
        if (self && (nq != nk || nq > ax::BQ)) return fail(h, FFB_ERR_UNSUPPORTED, "op_attention kind 6: needs nq == nk <= 128");
        if (ldq != H * 64 || ldk % 32 || ldk < H * 64) return fail(h, FFB_ERR_ARG, "op_attention kind 5/6: ldq must be H*64, ldk a multiple of 32");
        const size_t Mq = (size_t)G * nq, Rk = (size_t)G * nk;
        std::vector<int> seq_off(G + 1), row_off(G), vlen(G);
        for (int g = 0; g <= G; ++g) seq_off[g] = g;
        for (int g = 0; g < G; ++g) { row_off[g] = g * nk; vlen[g] = nk; }
        DevBuf qh, kh, vh, os, meta;
        int rc = FFB_OK;
        do {
            if (qh.ensure(2 * Mq * ldq * 2) != cudaSuccess || kh.ensure(2 * Rk * ldk * 2) != cudaSuccess || vh.ensure(2 * Rk * ldk * 2) != cudaSuccess ||
                os.ensure(2 * Mq * ldq * 2) != cudaSuccess || meta.ensure((3 * (size_t)G + 1) * sizeof(int)) != cudaSuccess) {
                rc = fail(h, FFB_ERR_CUDA, "op_attention: out of device memory"); break; }
            int* m = meta.as<int>();
            cudaMemcpyAsync(m, seq_off.data(), (G + 1) * sizeof(int), cudaMemcpyHostToDevice, s);
            cudaMemcpyAsync(m + G + 1, row_off.data(), G * sizeof(int), cudaMemcpyHostToDevice, s);
            cudaMemcpyAsync(m + 2 * G + 1, vlen.data(), G * sizeof(int), cudaMemcpyHostToDevice, s);
            if (cudaStreamSynchronize(s) != cudaSuccess) { rc = fail(h, FFB_ERR_CUDA, "op_attention: upload failed"); break; }
            // [rows, ld] contiguous arrays: split as one array, the parts are then rows*ld elements apart (the maps below say so)
            split_array_kernel<<<grid1d((long long)Mq * ldq / 4), 256, 0, s>>>(q, qh.as<uint16_t>(), (long long)Mq * ldq / 4, 1.0f, 2);
            split_array_kernel<<<grid1d((long long)Rk * ldk / 4), 256, 0, s>>>(k, kh.as<uint16_t>(), (long long)Rk * ldk / 4, 1.0f, 2);
            split_array_kernel<<<grid1d((long long)Rk * ldk / 4), 256, 0, s>>>(v, vh.as<uint16_t>(), (long long)Rk * ldk / 4, 1.0f, 2);
            h->launches += 3;
            CUtensorMap mq, mk, mv;
            if ((rc = encode_rows_map(h, &mq, qh.p, ldq, Mq, 32, ax::BQ, CU_TENSOR_MAP_SWIZZLE_64B)) != FFB_OK) break;
            if ((rc = encode_rows_map(h, &mk, kh.p, ldk, Rk, 32, ax::KC, CU_TENSOR_MAP_SWIZZLE_64B)) != FFB_OK) break;
            if ((rc = encode_rows_map(h, &mv, vh.p, ldk, Rk, 64, ax::KC, CU_TENSOR_MAP_SWIZZLE_128B)) != FFB_OK) break;
            ax::Params ap{};
            ap.Os = os.as<uint16_t>(); ap.os_stride = (long long)Mq * ldq; ap.ldo = ldq;
            if (self) { ap.mode = 1; ap.P = nq; ap.n_seqs = G; }
            else { ap.mode = 0; ap.seq_off = m; ap.q_mul = nq; ap.row_off = m + G + 1; ap.vlen = m + 2 * G + 1; ap.n_groups = G; }
            if ((rc = launch_attn_x(h, mq, mk, mv, ap, &seq_off, (double)G * nq * nk, PC_ATTN_TILED, nullptr, s)) != FFB_OK) break;
            sum_split_kernel<<<grid1d((long long)Mq * ldq), 256, 0, s>>>(os.as<uint16_t>(), (long long)Mq * ldq, out, (long long)Mq * ldq, 2);
            h->launches++;
            if (cudaStreamSynchronize(s) != cudaSuccess) { rc = fail(h, FFB_ERR_CUDA, "op_attention kind %d: %s", kind, cudaGetErrorString(cudaGetLastError())); break; }
        } while (0);
        qh.release(); kh.release(); vh.release(); os.release(); meta.release();
        return rc;
    }
    const int saved = h->opt_attn_mma, saved_fmt = h->tc_fmt;
    h->opt_attn_mma = (kind == 2) ? 1 : (kind == 3) ? 2 : 0;
    if (kind == 3) h->tc_fmt = 2;
    int rc;
    if (kind == 0) rc = launch_attn_rows(h, q, ldq, k, v, ldk, out, H * 64, G, nq, nk, nq, 0, nk, nq, nullptr, s);
    else {
        AttnGroups g{}; g.ragged = 0; g.nq = nq; g.nk = nk; g.q_stride = nq; g.q_off = 0; g.k_stride = nk; g.o_stride = nq;
        rc = launch_attn_tiled(h, q, ldq, k, v, ldk, out, H * 64, g, G, nq, (double)G * nq * nk, nullptr, s);
    }
    h->opt_attn_mma = saved; h->tc_fmt = saved_fmt;
    return rc;
}

}  // extern "C"
