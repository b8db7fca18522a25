/*
 * ffb200.h -- C ABI of libffb200.so: FaceFormer greedy pointer-decode on B200 (sm_100a).
 *
 * The reference (manycore-research/faceformer) is pure Python/PyTorch and has NO
 * FFI of its own; the boundary this library replaces is the model call
 *     Trainer.forward(batch) -> self.model(batch)        faceformer/trainer.py:27-28
 * i.e. SurfaceFormer_Parallel.forward_eval                faceformer/models/model_para.py:181-241
 * and  SurfaceFormer.forward_eval                         faceformer/models/model.py:169-219
 * over faceformer/transformer.py and faceformer/embedding.py.  The entry points
 * below are what a ctypes binding in faceformer/models/ would call (INTEGRATION.md
 * shows that binding).  Plain C types only: no torch / CUDA types in signatures
 * (a CUDA stream is passed as void*).
 *
 * Conventions
 *   - every function returns 0 on success, a negative ffb_status on failure;
 *     ffb_last_error() returns a message.  Nothing throws or exits across the ABI.
 *   - "loc" arguments say where caller buffers live: FFB_HOST or FFB_DEVICE.
 *     With FFB_HOST the library performs the H2D/D2H copies itself (on `stream`).
 *   - one handle per device per process; calls on one handle are not re-entrant;
 *     work is stream-ordered on the stream passed in.
 *   - the library owns only the handle, the packed weights and its workspaces
 *     (edge memory, cross-attention K/V cache, token buffer, activations).
 */
#ifndef FFB200_H
#define FFB200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FFB_ABI_VERSION 1

typedef struct ffb_handle ffb_handle;

typedef enum {
    FFB_OK = 0,
    FFB_ERR_ARG = -1,          /* bad argument / unsupported geometry            */
    FFB_ERR_CUDA = -2,         /* CUDA runtime error (message has the details)   */
    FFB_ERR_STATE = -3,        /* call sequence error (e.g. decode before encode)*/
    FFB_ERR_UNSUPPORTED = -4   /* valid in the reference, not supported here     */
} ffb_status;

enum { FFB_HOST = 0, FFB_DEVICE = 1 };
enum { FFB_MODE_PARALLEL = 0,  /* SurfaceFormer_Parallel, model_para.py          */
       FFB_MODE_SEQ2SEQ = 1 }; /* SurfaceFormer,          model.py               */

/* Constructor arguments of the reference models (model_para.py:14-19, model.py:14-18;
 * values from faceformer/config.py:27-49 and configs/\*.yml). */
typedef struct {
    int32_t abi_version;          /* FFB_ABI_VERSION                                         */
    int32_t mode;                 /* FFB_MODE_*                                              */
    int32_t num_model;            /* E, multiple of 64; head dim must be 64                  */
    int32_t num_head;             /* H = E / 64                                              */
    int32_t num_feedforward;      /* FF                                                      */
    int32_t num_encoder_layers;
    int32_t num_decoder_layers;
    int32_t in_dim;               /* num_points_per_line * point_dim (multiple of 4)         */
    int32_t num_lines;            /* max edges per wireframe; memory rows L = num_lines+num_token */
    int32_t num_token;            /* token.len (4)                                           */
    int32_t seq_len;              /* T = max_face_length (parallel) or label_seq_length      */
    int32_t device;               /* CUDA device ordinal                                     */
} ffb_config;

/* Options for ffb_set_option. */
enum {
    FFB_OPT_DEDUP_PAD = 1,   /* 1 (default): the F-n_i padded-anchor sequences of a wireframe
                                (all start with token 3, model_para.py:204-205) are decoded once
                                and broadcast; 0: decode all N*F sequences like the reference.   */
    FFB_OPT_PRUNE_LAST = 2   /* 1 (default in parallel mode): in the last decoder layer only the
                                last position is carried past self-attention (only pointer[-1]
                                is consumed, model_para.py:176); 0: all positions (seq2seq
                                default, so that 'pointer' [N,P,E] of model.py:217 is complete). */
};

/* Beam search (BASELINE.json configs[3], "beam=4"; parallel mode only).  The reference has NO beam search: the semantics are
 * specified by this build (oracle/beam_oracle.py) -- W hypotheses per anchor sequence scored by cumulative
 * log-softmax over the un-masked rows in float64, ties to the lowest (hypothesis, row) index, no per-sequence termination, the
 * reference's global stop predicate over all hypotheses.  value = W in [1, 8]; 1 (default) is exactly the greedy loop
 * (model_para.py:216-233).  `predict` then holds the best hypothesis of every anchor. */
enum { FFB_OPT_BEAM = 15 };

/* Precision of the two stages that dominate the pointer-logit error budget (both are ~1 % of a decode's FLOPs).
 * FFB_OPT_ENCODER_PRECISION: 2 (default) = embedding, encoder and the once-per-wireframe cross-attention K / V projections in
 *   float64 on the FP64 pipe (wireframes of <= 1024 memory rows; larger ones use mode 0); 0 = fp16x2 tcgen05 pipeline when the batch has
 *   >= 2048 memory rows (FFB_OPT_ENCODER_TC), else fp32 SIMT -- the throughput mode of the encoder-only workload (BASELINE configs[4]).
 * FFB_OPT_HEAD_FP64: 1 (default) = decoder.norm + project + the pointer dot product of the last position in float64; 0 = fp32. */
enum { FFB_OPT_ENCODER_PRECISION = 16, FFB_OPT_HEAD_FP64 = 17 };

/* Teacher-forced FORWARD pass of SurfaceFormer_Parallel.forward_train (faceformer/models/model_para.py:99-171 with scheduled_sampling_ratio = 0;
 * seq2seq handles: SurfaceFormer.forward_train, faceformer/models/model.py:98-157, with label / label_mask [N, T], i.e. label_rows = 1, F = 1):
 * encoder, then ONE decoder pass over the T - 1 input positions label[..., :-1] of every (wireframe, anchor slot) sequence with the causal mask
 * (model_para.py:72-74,120) and label_mask[..., :-1] as tgt_key_padding_mask (model_para.py:68-69,158-159), then project:
 * pointer [N * F, T - 1, E] fp32 with F = max(num_input) (the reference's outputs['pointer']).  label / label_mask: [N, label_rows, T] int64 / uint8
 * (non-zero = padding), label_rows >= F; only the first F rows of a wireframe are read (model_para.py:104).  Every teacher token must address an
 * un-masked memory row.  Forward only: there is no backward pass in this library -- this is the loss / accuracy evaluation of a teacher-forced
 * batch (faceformer/trainer.py:61-79), not a training step.  embedding (optional, may be NULL): [N, L, E] fp32 = the reference's
 * outputs['embedding'] before its replication per anchor slot (model_para.py:116,164) INCLUDING the rows of padded edges, which compute_loss's
 * softmax runs over (trainer.py:64-69); they come from a second, dense fp32 encoder pass (ffb_get_memory returns zeros in those rows).
 * The batch stays encoded WITHOUT padded-anchor de-duplication. */
int ffb_forward_train(ffb_handle* h, const float* coords, const uint8_t* pad_mask, const int64_t* num_input, int32_t N, const int64_t* label,
                      const uint8_t* label_mask, int32_t label_rows, float* pointer, float* embedding, int32_t loc, void* stream);

/* ---- one batch split over several GPUs (BASELINE.json configs[2]: "batch=128 sharded over 8xB200") -------------------------------
 * One process per GPU decodes a SHARE of the wireframes of one global batch.  forward_eval couples the wireframes of a batch in two
 * places only: F = max(num_input) (model_para.py:187) and the stop predicate all(next < 4) (model_para.py:232).  With
 *   FFB_OPT_FORCE_F = F of the global batch (set before ffb_encode; the share's predict is then [n_share, F, T]), and
 *   a connected stop exchange (every decode step ends with the ranks' local verdicts crossing NVLink through peer-mapped flag words;
 *   no host synchronisation, no collective call),
 * the shares concatenate to exactly the tensor a single GPU produces for the whole batch.  Protocol: every rank calls
 * ffb_stop_exchange_export (a 64-byte CUDA IPC handle of its flag buffer), the handles are all-gathered by the host framework
 * (torch.distributed), every rank calls ffb_stop_exchange_connect with all of them.  From then on every rank must call
 * ffb_decode_greedy the same number of times.  An fp16-range overflow cannot be re-run privately in this mode: it is reported as
 * FFB_ERR_UNSUPPORTED (use FFB_OPT_TC_FORMAT = 3 on every rank). */
enum { FFB_OPT_FORCE_F = 18 };
int ffb_stop_exchange_export(ffb_handle* h, void* ipc_handle_out /* 64 bytes */);
int ffb_stop_exchange_connect(ffb_handle* h, int32_t rank, int32_t world, const void* ipc_handles /* world x 64 bytes */);
int ffb_stop_exchange_disconnect(ffb_handle* h);

/* 1 (default): in the tensor-core encoder, self-attention over more than 256 keys per wireframe (or more than 255 wireframes) runs on
 * the tcgen05 key-streaming kernel (attn_l.cuh: 64-key blocks, exact two-pass softmax); 0 = the mma.sync kernel (attn_h.cuh). */
enum { FFB_OPT_ATTN_LONG = 20 };

/* 1 (default): fp16x2 GEMMs with at most num_SMs / 2 output tiles of 128 x 256 (small M: one wireframe per batch, seq2seq) use
 * 128 x 64 tiles instead: 4x as many CTAs stream the weights and the per-tile tensor time drops 4x.  0 = always 128 x 256. */
enum { FFB_OPT_SKINNY_GEMM = 21 };

/* 1 (default): in decoder layer 0 the q / k / v rows of prefix positions that already existed in the previous greedy step are taken from a
 * position-stable cache (their inputs -- memory[token], query_pos -- cannot change: exact, not the causal KV cache DESIGN.md rules out);
 * only the new position is normalised and projected.  Greedy loop on the half pipeline only (not with beams, not for forced prefixes). */
enum { FFB_OPT_L0_CACHE = 22 };

/* The whole greedy loop as ONE persistent cooperative kernel (csrc/persist.cuh): all decode steps in a single launch, grid-wide barriers
 * between the phases of a step, the stop predicate (model_para.py:232 / model.py:207-210) evaluated on the device.  0 = off; 1 (default) =
 * auto: batches of at most 896 decoder rows (sequences x (T - 1): one small wireframe per batch as in the reference's test loop,
 * trainer.py:51, and seq2seq) while no option forces one of the multi-kernel pipelines; 2 = wherever it is supported (fp16x2 operand format,
 * float64 head, beam width 1, no batch splitting, <= 64 wireframes); 3 = hybrid: the first steps of ANY supported batch, while sequences x prefix
 * length <= 896, then the per-step kernels take over from the same token / state buffers (meant for checkpoints whose decodes stop after a few
 * steps; on the synthetic batch-1 loop, where no decode stops early, it measured 2 % slower than mode 1 because the layer-0 cache is lost for the
 * rest of the decode).  Larger batches are tensor-bound and stay on the tcgen05 kernels. */
enum { FFB_OPT_PERSISTENT = 23 };

/* 1: ffb_encode computes the encoder memory only (embedding, encoder layers, final norm; ffb_get_memory reads it) -- no decode
 * workspaces are sized, the cross-attention K / V cache and the folded pointer head are skipped and ffb_decode_greedy is refused.
 * The encoder-only throughput workload of BASELINE.json configs[4] (2048-edge wireframes, batch 256). */
enum { FFB_OPT_ENCODE_ONLY = 19 };

/* Number of fp32 elements ffb_load_weights expects for this config: the reference
 * state_dict's float tensors, concatenated in state_dict order;
 * the two int64 `position` buffers are skipped.  Returns 0 for an invalid config. */
size_t ffb_weight_count(const ffb_config* cfg);

int ffb_create(const ffb_config* cfg, ffb_handle** out);
int ffb_destroy(ffb_handle* h);

/* Message for the last failure on this handle (h may be NULL: last ffb_create failure). */
const char* ffb_last_error(const ffb_handle* h);

int ffb_set_option(ffb_handle* h, int option, int value);

/* Replaces model.load_state_dict (main.py:46 via Trainer.load_from_checkpoint). */
int ffb_load_weights(ffb_handle* h, const float* blob, size_t count, int loc, void* stream);

/* Everything before the decode loop (model_para.py:183-214 / model.py:171-186):
 * embedding, encoder, anchors, and -- once per wireframe instead of once per step --
 * the cross-attention K/V projections of every decoder layer.
 *   coords    float [N, num_lines, in_dim]     inputs['input'] flattened over (P,D)
 *   pad_mask  uint8 [N, num_lines]             inputs['input_mask'] (1 = padding); must be of
 *                                              prefix form (valid edges first), as
 *                                              data_para.py:67-68 builds it, else FFB_ERR_UNSUPPORTED
 *   num_input int64 [N]                        inputs['num_input'] (parallel mode; NULL in seq2seq)
 * pad_mask and num_input are read on the host (the reference also syncs on max(num_input),
 * model_para.py:187); with loc == FFB_DEVICE they are copied back first. */
int ffb_encode(ffb_handle* h, const float* coords, const uint8_t* pad_mask,
               const int64_t* num_input, int32_t n_wireframes, int loc, void* stream);

/* Geometry of the encoded batch: F = max(num_input) (1 in seq2seq), B = N*F sequences the
 * reference would decode, B_eff = sequences actually decoded, R = packed memory rows. */
int ffb_batch_info(const ffb_handle* h, int32_t* n_wireframes, int32_t* F, int64_t* B,
                   int64_t* B_eff, int64_t* R);

/* The greedy loop (model_para.py:216-240 / model.py:193-218), entirely on the device:
 * no host synchronisation per step; early stop is a device-side flag.
 *   predict   int64 [N, F, T] (parallel) or [N, T] (seq2seq)      inputs['predict']
 *   steps_run number of executed decode steps S (host int32, written after a stream sync;
 *             may be NULL, in which case the call is fully asynchronous for loc == FFB_DEVICE) */
int ffb_decode_greedy(ffb_handle* h, int64_t* predict, int loc, int32_t* steps_run, void* stream);

/* ffb_encode + ffb_decode_greedy: the whole `model(batch)` call. */
int ffb_forward_eval(ffb_handle* h, const float* coords, const uint8_t* pad_mask,
                     const int64_t* num_input, int32_t n_wireframes, int64_t* predict,
                     int loc, int32_t* steps_run, void* stream);

/* Input featurisation, the step right before the path: what ABCDataset_Parallel.__getitem__ does per wireframe
 * with sample_points (faceformer/datasets/data_para.py:8-25,59-68), for a whole batch on the device.
 *   points       double [n_points_total, 2]   all edge polylines of the batch, concatenated (JSON coordinates are doubles)
 *   edge_off     int64  [n_edges_total + 1]   prefix offsets of the edges into `points` (an edge needs >= 1 point)
 *   wf_edge_off  int64  [N + 1]               prefix offsets of the wireframes into the edges (<= num_lines edges each)
 * outputs (same location as the inputs):
 *   coords       float  [N, num_lines, in_dim] 2-point edges resampled on the segment (float64 arithmetic like numpy, rounded to
 *                                              float32), longer polylines index-resampled; padded slots zero     inputs['input']
 *   pad_mask     uint8  [N, num_lines]         1 = padding                                                       inputs['input_mask']
 *   num_input    int64  [N]                                                                                      inputs['num_input']
 * Requires point_dim 2 (in_dim = 2 * num_points_per_line).  Results are bit-identical to the reference's. */
int ffb_featurize(ffb_handle* h, const double* points, const int64_t* edge_off, const int64_t* wf_edge_off, int32_t n_wireframes,
                  float* coords, uint8_t* pad_mask, int64_t* num_input, int loc, void* stream);

/* Prediction parsing, the step right after the path: for every predicted sequence of `predict` [N, F, T] what
 * Trainer.parse_parallel_faces (faceformer/trainer.py:196-206, predict half) and filter_faces_by_encloseness
 * (faceformer/post_processing.py:8-20 over dataset/tests/check_faces_enclosed.py:11-46) do in Python, one thread per sequence:
 * cut after the first face-type token, drop the token offset and out-of-range indices, require every edge to start where the previous
 * one ended (|dx| < tol and |dy| < tol on the polylines' end points) and every loop to close, roll each loop so that its smallest
 * index comes first and order the loops by their first index.  `points` / `edge_off` / `wf_edge_off` as in ffb_featurize.
 *   valid      uint8 [N, F]     1 = the sequence yields a face (check_enclosed: ... an enclosed one)
 *   face_type  int32 [N, F]     type token - face_type_offset (of the cut position; meaningful where valid)
 *   n_loops    int32 [N, F]     loops of the face (0 when check_enclosed == 0)
 *   loop_len   int32 [N, F, T]  first n_loops entries: loop lengths in canonical order
 *   indices    int32 [N, F, T]  first n_indices entries: edge indices, loops concatenated in canonical order
 *   n_indices  int32 [N, F]
 * check_enclosed = post_process.is_coedge (config.py:52), tol = post_process.enclosedness_tol (config.py:51: 2e-4). */
int ffb_parse_faces(ffb_handle* h, const int64_t* predict, int32_t n_wireframes, int32_t F, const double* points,
                    const int64_t* edge_off, const int64_t* wf_edge_off, double tol, int32_t check_enclosed, uint8_t* valid,
                    int32_t* face_type, int32_t* n_loops, int32_t* loop_len, int32_t* indices, int32_t* n_indices, int loc, void* stream);

/* Parity hooks (used by tests; they do not change decode results). */

/* Encoder memory, float [N, L, E] (= inputs['embedding'] of model.py:216); rows of padded
 * edges are written as zeros (the reference computes values there that nothing reads). */
int ffb_get_memory(ffb_handle* h, float* memory, int loc, void* stream);

/* Error-budget hook: REPLACE the encoder memory of the encoded batch by caller data, float [N, L, E] (rows of padded edges are
 * ignored), and recompute the cross-attention K / V cache from it.  Feeding the reference's float64 memory isolates the
 * decoder + pointer-head contribution to the logit error (profiles/logit_noise.py). */
int ffb_set_memory(ffb_handle* h, const float* memory, int loc, void* stream);

/* Masked pointer logits [B, L] (select_next up to masked_fill, model_para.py:173-177) of the
 * LAST executed decode step, expanded to the reference's B = N*F rows. */
int ffb_get_last_logits(ffb_handle* h, float* logits, int loc, void* stream);

/* seq2seq only: inputs['pointer'] [N, P, E] of the last executed step (model.py:217);
 * requires FFB_OPT_PRUNE_LAST = 0.  *P_out receives P. */
int ffb_get_last_pointer(ffb_handle* h, float* pointer, int32_t* P_out, int loc, void* stream);

/* One loop body on a caller-supplied token prefix (forced-prefix parity):
 *   prefix int64 [P, B] with B = N*F (parallel) or N (seq2seq), logits float [B, L].
 * Requires FFB_OPT_DEDUP_PAD = 0 at encode time in parallel mode. */
int ffb_forced_prefix_logits(ffb_handle* h, const int64_t* prefix, int32_t P, float* logits,
                             int loc, void* stream);

/* Number of decodes that had to be re-run in bf16x3 because an activation left the fp16 range (see FFB_OPT_TC_FORMAT). */
int ffb_fp16_fallbacks(const ffb_handle* h);

/* After a decode with FFB_OPT_BEAM = W > 1: all hypotheses, best first.
 *   beams  int64  [N, F, W, T]    scores double [N, F, W] (cumulative log-probability; -inf = dead hypothesis) */
int ffb_get_beams(ffb_handle* h, int64_t* beams, double* scores, int loc, void* stream);

/* A fully asynchronous ffb_decode_greedy (device buffers, steps_run == NULL) cannot look at the fp16-range flag of the default
 * operand format and therefore cannot re-run by itself.  After such a call, ffb_overflowed() synchronises `stream` and reports in
 * *overflowed whether an activation left the fp16 range.  If so the predictions of that decode are INVALID: the handle has been
 * switched to the bf16x3 format (sticky, ffb_fp16_fallbacks() is incremented) and the batch must be encoded and decoded again.
 * The syncing forms of ffb_decode_greedy (steps_run != NULL or host buffers) do all of this internally. */
int ffb_overflowed(ffb_handle* h, int32_t* overflowed, void* stream);

/* Decode steps whose kernels the last ffb_decode_greedy launched: the host watches a pinned mirror of the device-side stop flag and
 * stops launching once the early-stop predicate (model_para.py:232 / model.py:207-210) has fired; == executed steps + the few
 * steps that were already queued when the flag arrived. */
int ffb_steps_launched(const ffb_handle* h);

/* 1 if the last ffb_decode_greedy ran the whole loop inside the persistent cooperative kernel (FFB_OPT_PERSISTENT), else 0. */
int ffb_used_persistent(const ffb_handle* h);

/* Count of this library's kernels launched on the handle since creation (bench `gpu_launches`). */
int64_t ffb_kernel_launches(const ffb_handle* h);

/* Device-time of the phases of the last forward (ms, CUDA events; valid after a stream sync):
 * out[0] = encode (embedding + encoder + cross-K/V), out[1] = decode loop.  Timing is only
 * recorded when enabled with ffb_set_option(h, FFB_OPT_TIMING, 1). */
enum { FFB_OPT_TIMING = 3 };
int ffb_phase_times(ffb_handle* h, float* out_ms, int32_t n);

/* Per-kernel-class device time (CUDA events recorded around EVERY launch on the launching stream) and
 * algorithmic FLOPs since profiling was enabled with ffb_set_option(h, FFB_OPT_PROFILE, 1) or last read.
 * Classes: 0 linear (fp32 SIMT), 1 layernorm, 2 attention (decoder self), 3 attention (encoder / cross),
 * 4 pointer, 5 other, 6 linear (tcgen05 split-precision tensor-core GEMM).  Perturbs timing slightly: bench.py uses it on an extra, untimed step.  Synchronises the device. */
enum { FFB_OPT_PROFILE = 4, FFB_PROFILE_CLASSES = 7 };
int ffb_profile_read(ffb_handle* h, int32_t n_classes, float* ms, double* flops, int64_t* launches);

/* Tensor-core path selection for the decode-step linear layers (gemm_tc.cuh: split-precision tcgen05 GEMM, operand format per
 * FFB_OPT_TC_FORMAT): 0 = off (fp32 SIMT everywhere), 1 = auto (default; since round 2 every step of a geometry on the 256 grid, the
 * tcgen05 GEMM being faster than the SIMT kernel at every M), 2 = force (same as auto today; kept for tests). */
enum { FFB_OPT_TENSOR_CORE = 5 };
/* Attention core of the decode loop where the tcgen05 kernels (FFB_OPT_ATTN_X) do not apply: 2 (default) = fp16x2 operands (the half
 * pipeline: attn_h.cuh / attn_f16.cuh on mma.sync m16n8k16; same overflow fallback as the GEMM format), 1 = mma.sync m16n8k8 3xTF32
 * split (attn_mma.cuh; also the fp32 encoder's), 0 = fp32 SIMT kernels.  Values other than 2 switch the half pipeline off. */
enum { FFB_OPT_ATTN_MMA = 6 };
/* Operand format of the tensor-core GEMM: 2 (default) = fp16x2 (3 MMA passes; weights pre-scaled by a power of two;
 * if an activation exceeds the fp16 range the decode is transparently re-run in format 3 and the handle stays there),
 * 3 = bf16x3 (6 MMA passes, full fp32 range). */
enum { FFB_OPT_TC_FORMAT = 7 };
/* 1: de-phase the persistent GEMM CTAs (default 0: measured no effect); so that their store bursts overlap other CTAs' mainloops; 0 = off. */
enum { FFB_OPT_STAGGER = 8 };
/* 1 (default): fp16x2 GEMM writes fp32 outputs with asynchronous TMA bulk stores (reduce-add for the in-place residual);
 * 0 = coalesced st.global epilogue. */
enum { FFB_OPT_TMA_EPILOGUE = 9 };
/* Attention cores of the fp16x2 pipeline on the tcgen05 kernel (attn_x.cuh), bit mask: 1 = decoder cross-attention (needs <= 256
 * keys per wireframe and <= 255 wireframes per batch, otherwise the mma.sync kernel is used automatically), 2 = decoder
 * self-attention (prefix <= 128).  Default 3; 0 = always the mma.sync kernel. */
enum { FFB_OPT_ATTN_X = 10 };
/* fp16x2 GEMM pipeline variant: 0 = 4 operand stages + 1 epilogue staging buffer per warp, 1 = 3 stages + 2 staging buffers,
 * 2 (default) = variant 1 for plain / split stores and variant 0 for the in-place residual,
 * 3 = experimental CTA-pair kernel (tcgen05 cta_group::2, gemm_tc2.cuh; correct but slower in round 1). */
enum { FFB_OPT_GEMM_VARIANT = 11 };
/* 1 (default): encoder layers and the once-per-wireframe cross-attention K / V projections run on the tcgen05 pipeline (fp16x2 GEMMs,
 * tcgen05 attention) when the batch has >= 2048 memory rows and <= 256 rows per wireframe; 0 = always fp32 SIMT + 3xTF32 mma.sync. */
enum { FFB_OPT_ENCODER_TC = 12 };
/* Programmatic dependent launch (programmatic stream serialization) for the decode-step kernels: each kernel's CTAs are scheduled as
 * its predecessor's retire and wait (griddepcontrol.wait) before touching memory.  0 = off, 1 = on, 2 (default) = on for batches of
 * <= 32768 decode rows (launch-latency bound: measured +5..10 % at one wireframe per batch, -3 % on the 32-wireframe bench batch). */
enum { FFB_OPT_PDL = 13 };
/* 1 (default): the pointer head (select_next) streams a wireframe's memory rows once per 16 of its sequences
 * (pointer_batched_kernel; logits and tokens bit-identical to the per-sequence kernel); 0 = one CTA per sequence. */
enum { FFB_OPT_POINTER_BATCHED = 14 };

/* ---- op-level test hooks: run ONE kernel of the path on caller data (device pointers). ----
 * They exist so that tests can compare each kernel with the oracle's primitive. */

/* C[M,N] = act((A (+pos[r % pos_mod] on columns n < pos_cols)) W^T + bias) (+ R)          */
int ffb_op_linear(ffb_handle* h, const float* A, const float* W, const float* bias, const float* R,
                  const float* pos, int32_t pos_mod, int32_t pos_cols, float* C,
                  int32_t M, int32_t N, int32_t K, int32_t relu, void* stream);
/* y = LayerNorm(x) over rows of length E (eps 1e-5)                                         */
int ffb_op_layernorm(ffb_handle* h, const float* x, const float* gamma, const float* beta,
                     float* y, int32_t M, int32_t E, void* stream);
/* Multi-head attention core (softmax(q k^T / 8) v) for G equal-sized groups:
 * q [G*nq, ldq], k/v [G*nk, ldk] with head h in columns [64h, 64h+64); out [G*nq, H*64].
 * kind 0 = SIMT warp-per-row kernel, 1 = SIMT tiled kernel, 2 = 3xTF32 mma.sync kernel, 3 = fp16x2 mma.sync kernel, 4 = fp16x2 kernel
 * with pre-split inputs (attn_h.cuh; the hook splits q/k/v first; q/k/v must be contiguous [rows, ld] arrays), 5 / 6 = tcgen05 kernel
 * (attn_x.cuh) in CROSS / SELF mode, 7 = tcgen05 key-streaming kernel (attn_l.cuh; nq == nk, any number of keys).  */
int ffb_op_attention(ffb_handle* h, int32_t kind, const float* q, int32_t ldq, const float* k,
                     const float* v, int32_t ldk, float* out, int32_t G, int32_t nq, int32_t nk,
                     int32_t H, void* stream);

/* The tensor-core GEMM alone on fp32 caller data (device pointers): A [M,K] and W [N,K] are split to bf16x3 inside,
 * C[M,N] = act(A W^T + bias) (+ R).  via_split != 0 routes the result through the kernel's bf16x3 output format
 * (used for the FFN hidden activations) and re-sums it.  N %% 256 == 0, K %% 32 == 0. */
int ffb_op_linear_tc(ffb_handle* h, const float* A, const float* W, const float* bias, const float* R, float* C,
                     int32_t M, int32_t N, int32_t K, int32_t relu, int32_t via_split, void* stream);

/* Tuning aid: average device time (ms, CUDA events) of `iters` launches of the tensor-core GEMM on zero-filled
 * operands of the given shape.  flags: 1 bias, 2 residual, 4 relu, 8 bf16x3 split output, 16 no output stores. */
int ffb_bench_linear_tc(ffb_handle* h, int32_t M, int32_t N, int32_t K, int32_t flags, int32_t iters, float* ms_out, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* FFB200_H */
