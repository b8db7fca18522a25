"""CPU oracle for FaceFormer's greedy pointer-decode path -- TEST INFRASTRUCTURE ONLY.

This file is a plain numpy fp32 restatement of the reference's evaluation path.
It is the *checker*: only ``tests/``, ``__graft_entry__.smoke()`` and
``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may import it.  The
product (``faceformer_b200``) never imports it and has no CPU fallback.

It restates the reference AS WRITTEN, including its waste (memory replicated F
times and cross-attention K/V re-projected at every step), because it also
stands in for the reference's CPU cost on machines where /root/reference is
absent.

Parity status: PINNED against outputs of the reference itself.  The reference
ships no tests, golden vectors or fixtures for this path,
so ``oracle/make_golden.py`` imports the unmodified reference model from
/root/reference, runs ``forward_eval`` on seeded synthetic weights/inputs
(``faceformer_b200.synth``) and commits the outputs under ``tests/golden/``;
``tests/test_oracle_golden.py`` checks this restatement against them (tokens
exact, logits/memory within 1e-4).

The arithmetic itself lives in PyTorch (un-vendored third-party; pinned
``pytorch=1.7.1`` in /root/reference/environment.yml:8, torch 2.11.0 installed
here).  ``multi_head_attention`` below restates the published algorithm of
``torch.nn.functional.multi_head_attention_forward`` (torch 2.11
functional.py:6244-6696; same operation sequence in 1.7.1).

Reference citations are ``file:line`` relative to /root/reference/.
"""
from __future__ import annotations

import numpy as np

F32 = np.float32
LN_EPS = F32(1e-5)                               # nn.LayerNorm default eps
F32_MIN = np.finfo(np.float32).min               # faceformer/utils.py:16-20

MODE_PARALLEL = 0
MODE_SEQ2SEQ = 1


# ----------------------------------------------------------------------------- primitives
def linear(x, w, b=None):
    """nn.Linear: y = x W^T + b."""
    y = np.matmul(x.reshape(-1, x.shape[-1]), w.T)              # one sgemm, not a batch of small ones
    if b is not None:
        y = y + b
    return y.reshape(x.shape[:-1] + (w.shape[0],)).astype(F32, copy=False)


def layer_norm(x, w, b):
    """nn.LayerNorm over the last dim, biased variance, eps=1e-5
    (transformer.py:138-139,199-201; model_para.py:37,43)."""
    mean = x.mean(axis=-1, keepdims=True, dtype=F32)
    xc = x - mean
    var = (xc * xc).mean(axis=-1, keepdims=True, dtype=F32)
    return (xc / np.sqrt(var + LN_EPS) * w + b).astype(F32, copy=False)


def softmax_lastdim(x):
    m = x.max(axis=-1, keepdims=True)
    e = np.exp(x - m, dtype=F32)
    return (e / e.sum(axis=-1, keepdims=True, dtype=F32)).astype(F32, copy=False)


def multi_head_attention(query, key, value, in_w, in_b, out_w, out_b, num_head,
                         key_padding_mask=None, attn_mask=None):
    """torch F.multi_head_attention_forward, eval mode, separate q/k/v tensors.

    query [Lq,B,E], key/value [Lk,B,E]; key_padding_mask bool [B,Lk] (True = ignore).
    Sequence (torch 2.11 functional.py): packed in-proj via w.chunk(3) (5866-5873);
    bool mask -> 0/-inf float (6214-6217); q scaled by sqrt(1/d_head) BEFORE QK^T
    (6632); baddbmm(mask, q, k^T) (6637-6642); softmax(-1) (6643); bmm(attn, v)
    (6647); out_proj (6653).  Called from transformer.py:169-171,244-251.
    """
    Lq, B, E = query.shape
    Lk = key.shape[0]
    H = num_head
    d = E // H
    wq, wk, wv = in_w[:E], in_w[E:2 * E], in_w[2 * E:]
    bq, bk, bv = in_b[:E], in_b[E:2 * E], in_b[2 * E:]
    q = linear(query, wq, bq)
    k = linear(key, wk, bk)
    v = linear(value, wv, bv)
    q = q.reshape(Lq, B * H, d).transpose(1, 0, 2)
    k = k.reshape(Lk, B * H, d).transpose(1, 0, 2)
    v = v.reshape(Lk, B * H, d).transpose(1, 0, 2)
    q = q * F32(np.sqrt(1.0 / float(d)))
    s = np.matmul(q, k.transpose(0, 2, 1))                               # [B*H, Lq, Lk]
    if attn_mask is not None:                                      # bool [Lq,Lk], True = ignore
        s = s + np.where(attn_mask, F32(-np.inf), F32(0))[None]
    if key_padding_mask is not None:
        m = np.where(key_padding_mask, F32(-np.inf), F32(0)).astype(F32)   # [B,Lk]
        m = np.repeat(m[:, None, None, :], H, axis=1).reshape(B * H, 1, Lk)
        s = s + m
    p = softmax_lastdim(s.astype(F32, copy=False))
    o = np.matmul(p, v)                                                  # [B*H, Lq, d]
    o = o.transpose(1, 0, 2).reshape(Lq * B, E)
    o = linear(o, out_w, out_b)
    return o.reshape(Lq, B, E)


# ----------------------------------------------------------------------------- model pieces
def vanilla_embedding(sd, coord):
    """VanillaEmedding.forward (embedding.py:23-38): coord [N,L,P,D] -> [N,4+L,E]."""
    N = coord.shape[0]
    tok = sd["val_enc.embedding_token.weight"]
    token_embed = np.broadcast_to(tok[None], (N,) + tok.shape)
    x = coord.reshape(coord.shape[0], coord.shape[1], -1)          # embed_points: flatten(-2,-1)
    h = np.maximum(linear(x, sd["val_enc.embedding_value.0.weight"], sd["val_enc.embedding_value.0.bias"]), 0)
    c = linear(h, sd["val_enc.embedding_value.2.weight"], sd["val_enc.embedding_value.2.bias"])
    return np.concatenate([token_embed, c], axis=1).astype(F32)


def position_embedding(sd, prefix, length):
    """PositionEmbeddingLearned.forward (embedding.py:106-108): first `length` table rows, [1,len,E]."""
    return sd[prefix + ".pos_embed.weight"][sd[prefix + ".position"][0, :length]][None]


def process_masks(input_mask, num_token=4):
    """model_para.py:62-70 / model.py:61-69: prepend `num_token` un-masked columns."""
    pad = np.zeros((input_mask.shape[0], num_token), dtype=bool)
    return np.concatenate([pad, input_mask.astype(bool)], axis=1)


def encoder_layer_pre(sd, p, H, src, key_padding_mask, pos):
    """TransformerEncoderLayer.forward_pre (transformer.py:164-176)."""
    src2 = layer_norm(src, sd[p + ".norm1.weight"], sd[p + ".norm1.bias"])
    q = k = src2 + pos
    src2 = multi_head_attention(q, k, src2, sd[p + ".self_attn.in_proj_weight"], sd[p + ".self_attn.in_proj_bias"],
                                sd[p + ".self_attn.out_proj.weight"], sd[p + ".self_attn.out_proj.bias"], H,
                                key_padding_mask=key_padding_mask)
    src = src + src2
    src2 = layer_norm(src, sd[p + ".norm2.weight"], sd[p + ".norm2.bias"])
    src2 = linear(np.maximum(linear(src2, sd[p + ".linear1.weight"], sd[p + ".linear1.bias"]), 0),
                  sd[p + ".linear2.weight"], sd[p + ".linear2.bias"])
    return (src + src2).astype(F32, copy=False)


def encoder(sd, cfg, src, key_padding_mask, pos):
    """TransformerEncoder.forward (transformer.py:70-83) with final norm (model_para.py:37-38)."""
    out = src
    for l in range(cfg["num_encoder_layers"]):
        out = encoder_layer_pre(sd, f"encoder.layers.{l}", cfg["num_head"], out, key_padding_mask, pos)
    return layer_norm(out, sd["encoder.norm.weight"], sd["encoder.norm.bias"])


def decoder_layer_pre(sd, p, H, tgt, memory, memory_key_padding_mask, pos, query_pos, tgt_mask=None, tgt_key_padding_mask=None):
    """TransformerDecoderLayer.forward_pre (transformer.py:235-256)."""
    tgt2 = layer_norm(tgt, sd[p + ".norm1.weight"], sd[p + ".norm1.bias"])
    q = k = tgt2 + query_pos
    tgt2 = multi_head_attention(q, k, tgt2, sd[p + ".self_attn.in_proj_weight"], sd[p + ".self_attn.in_proj_bias"],
                                sd[p + ".self_attn.out_proj.weight"], sd[p + ".self_attn.out_proj.bias"], H,
                                attn_mask=tgt_mask, key_padding_mask=tgt_key_padding_mask)
    tgt = tgt + tgt2
    tgt2 = layer_norm(tgt, sd[p + ".norm2.weight"], sd[p + ".norm2.bias"])
    tgt2 = multi_head_attention(tgt2 + query_pos, memory + pos, memory,
                                sd[p + ".multihead_attn.in_proj_weight"], sd[p + ".multihead_attn.in_proj_bias"],
                                sd[p + ".multihead_attn.out_proj.weight"], sd[p + ".multihead_attn.out_proj.bias"], H,
                                key_padding_mask=memory_key_padding_mask)
    tgt = tgt + tgt2
    tgt2 = layer_norm(tgt, sd[p + ".norm3.weight"], sd[p + ".norm3.bias"])
    tgt2 = linear(np.maximum(linear(tgt2, sd[p + ".linear1.weight"], sd[p + ".linear1.bias"]), 0),
                  sd[p + ".linear2.weight"], sd[p + ".linear2.bias"])
    return (tgt + tgt2).astype(F32, copy=False)


def decoder(sd, cfg, tgt, memory, memory_key_padding_mask, pos, query_pos, tgt_mask=None, tgt_key_padding_mask=None):
    """TransformerDecoder.forward (transformer.py:95-124) incl. final norm (115-116)."""
    out = tgt
    for l in range(cfg["num_decoder_layers"]):
        out = decoder_layer_pre(sd, f"decoder.layers.{l}", cfg["num_head"], out, memory,
                                memory_key_padding_mask, pos, query_pos, tgt_mask, tgt_key_padding_mask)
    return layer_norm(out, sd["decoder.norm.weight"], sd["decoder.norm.bias"])


def gather_rows(memory, predicts):
    """torch.gather(memory, 0, predicts[...,None].repeat(1,1,E)) (model_para.py:217-219):
    tgt[p,b,:] = memory[predicts[p,b], b, :]."""
    B = memory.shape[1]
    return memory[predicts, np.arange(B)[None, :], :]


def pointer_logits(memory, pointer_last, input_mask):
    """select_next up to the masked logits (model_para.py:173-177): [B,L]."""
    emb = memory.transpose(1, 0, 2)                                # [B,L,E]
    logit = np.matmul(emb, pointer_last[:, :, None])[..., 0]             # bmm -> [B,L]
    return np.where(input_mask, F32_MIN, logit).astype(F32)


def decode_step(sd, cfg, memory, input_mask, pos, query_pos, predicts):
    """One iteration body of the greedy loop (model_para.py:217-227 / model.py:196-203).

    memory [L,B,E], input_mask [B,L], predicts [P,B] int64.
    Returns (next_token [B] int64, logits [B,L] f32, pointer [P,B,E])."""
    P = predicts.shape[0]
    tgt = gather_rows(memory, predicts)
    ptr = decoder(sd, cfg, tgt, memory, input_mask, pos, query_pos[:P])        # NO tgt_mask in eval
    ptr = linear(ptr, sd["project.weight"], sd["project.bias"])               # all P positions (model_para.py:225)
    logits = pointer_logits(memory, ptr[-1], input_mask)
    return np.argmax(logits, axis=1).astype(np.int64), logits, ptr             # first max index


def encode(sd, cfg, mode, inputs):
    """Everything before the loop (model_para.py:183-214 / model.py:171-186).
    Returns memory [L,N,E], input_mask [N,L], pos [L,1,E], query_pos [T,1,E]."""
    coord = np.asarray(inputs["input"], dtype=F32)
    input_mask = process_masks(np.asarray(inputs["input_mask"]), cfg["num_token"])
    val = vanilla_embedding(sd, coord)                                         # [N,L,E]
    pos = position_embedding(sd, "pos_enc", val.shape[1])                      # [1,L,E]
    T = inputs["label"].shape[-1]      # parallel: label.transpose(1,2).size(1) == T; seq2seq: label.size(1) == T
    qpos = position_embedding(sd, "query_pos_enc", T)                          # [1,T,E]
    source = val.transpose(1, 0, 2)
    pos = pos.transpose(1, 0, 2)
    qpos = qpos.transpose(1, 0, 2)
    memory = encoder(sd, cfg, source, input_mask, pos)
    return memory, input_mask, pos, qpos


def forward_eval(sd, cfg, mode, inputs, return_trace=False, max_steps=None):
    """SurfaceFormer_Parallel.forward_eval (model_para.py:181-241) or
    SurfaceFormer.forward_eval (model.py:169-219).

    cfg is a plain dict (ModelConfig.to_dict()).  Returns dict with 'predict'
    (int64 [N,F,T] or [N,T]), 'steps' (executed decode steps S) and, when
    return_trace, per-step logits and the encoder memory.  max_steps (test aid)
    truncates the loop: the first max_steps+1 token rows are then still exact."""
    memory, input_mask, pos, qpos = encode(sd, cfg, mode, inputs)
    N = memory.shape[1]
    num_token = cfg["num_token"]
    trace = []
    if mode == MODE_PARALLEL:
        T = cfg["max_face_length"]
        num_input = np.asarray(inputs["num_input"]).astype(np.int64)
        Fm = int(num_input.max())                                              # model_para.py:187
        anchors = np.tile(np.arange(Fm, dtype=np.int64), (1, N, 1))            # model_para.py:201 (NOT +4)
        for i, ne in enumerate(num_input):
            anchors[:, i, int(ne):] = num_token - 1                            # model_para.py:204-205
        predicts = anchors.reshape(1, N * Fm)
        mem_rep = np.repeat(memory, Fm, axis=1)                                # model_para.py:212
        mask_rep = np.repeat(input_mask, Fm, axis=0)                           # model_para.py:214
        steps = 0
        for step in range(T - 1 if max_steps is None else min(T - 1, max_steps)):
            nxt, logits, _ = decode_step(sd, cfg, mem_rep, mask_rep, pos, qpos, predicts)
            predicts = np.concatenate([predicts, nxt[None]], axis=0)
            steps += 1
            if return_trace:
                trace.append(logits)
            if np.all(nxt < num_token):                                        # model_para.py:232
                break
        predicts = np.concatenate([predicts, np.zeros((T - predicts.shape[0], predicts.shape[1]), np.int64)], 0)
        out = {"predict": predicts.T.reshape(N, Fm, T), "steps": steps}
    else:
        T = cfg["label_seq_length"]
        SOS, EOS = 1, 3                                                        # config.py:42-44
        predicts = np.full((1, N), SOS, dtype=np.int64)                        # model.py:190
        eos_found = 0
        steps = 0
        ptr = None
        for step in range(T - 1 if max_steps is None else min(T - 1, max_steps)):
            nxt, logits, ptr = decode_step(sd, cfg, memory, input_mask, pos, qpos, predicts)
            predicts = np.concatenate([predicts, nxt[None]], axis=0)
            steps += 1
            if return_trace:
                trace.append(logits)
            eos_found += int((nxt == EOS).sum())                               # model.py:207 (cumulative)
            if eos_found == N:                                                 # model.py:209
                break
        predicts = np.concatenate([predicts, np.zeros((T - predicts.shape[0], predicts.shape[1]), np.int64)], 0)
        out = {"predict": predicts.T.copy(), "steps": steps,
               "embedding": memory.transpose(1, 0, 2), "pointer": ptr.transpose(1, 0, 2)}
    if return_trace:
        out["logits"] = trace
        out["memory"] = memory
    return out


def forced_prefix_logits(sd, cfg, mode, inputs, prefix):
    """Pointer logits for an arbitrary token prefix (mitigation i).

    prefix: int64 [P,B] with B = N*F (parallel, F = max(num_input)) or N (seq2seq).
    Returns logits [B,L] exactly as the loop body would compute them."""
    memory, input_mask, pos, qpos = encode(sd, cfg, mode, inputs)
    if mode == MODE_PARALLEL:
        Fm = int(np.asarray(inputs["num_input"]).max())
        memory = np.repeat(memory, Fm, axis=1)
        input_mask = np.repeat(input_mask, Fm, axis=0)
    _, logits, _ = decode_step(sd, cfg, memory, input_mask, pos, qpos, np.asarray(prefix, dtype=np.int64))
    return logits


def forward_train(sd, cfg, inputs, mode=MODE_PARALLEL):
    """SurfaceFormer_Parallel.forward_train (model_para.py:99-171) / SurfaceFormer.forward_train (model.py:98-157) with
    scheduled_sampling_ratio = 0: the teacher-forced pass.

    Returns the reference's outputs: embedding [B, L, E], pointer [B, T-1, E], label [B, T-1] with B = N*F (parallel, model_para.py:164-166)
    or B = N (seq2seq, model.py:154-156)."""
    label, label_mask = np.asarray(inputs["label"]), np.asarray(inputs["label_mask"], bool)
    F = 1
    if mode == MODE_PARALLEL:
        F = int(np.max(inputs["num_input"]))                                    # model_para.py:103
        label, label_mask = label[:, :F, :], label_mask[:, :F, :]              # :104
    else:
        label, label_mask = label[:, None, :], label_mask[:, None, :]          # one sequence per wireframe
    tgt_kpm = label_mask[..., :-1]                                             # process_masks: "tgt is 1 shorter" (:68-69)
    memory, input_mask, pos, qpos = encode(sd, cfg, mode, inputs)              # :107-116 (memory [L,N,E], qpos [T,1,E])
    tgt = label.transpose(2, 0, 1)                                             # patch_target (:82-88): [T,N,F]
    target, lab = tgt[:-1], tgt[1:]
    qpos = qpos[:-1]
    N = memory.shape[1]
    T1 = target.shape[0]
    memory = np.repeat(memory, F, axis=1)                                      # :116 repeat_interleave -> [L, N*F, E]
    causal = np.triu(np.ones((T1, T1), bool), k=1)                             # generate_square_subsequent_mask (:72-74): True = ignore
    tgt_emb = gather_rows(memory, target.reshape(T1, N * F))                   # :146-151
    ptr = decoder(sd, cfg, tgt_emb, memory, np.repeat(input_mask, F, axis=0), pos, qpos, tgt_mask=causal,
                  tgt_key_padding_mask=tgt_kpm.reshape(N * F, T1))             # :158-159
    ptr = linear(ptr, sd["project.weight"], sd["project.bias"])               # :161
    return dict(embedding=memory.transpose(1, 0, 2), pointer=ptr.transpose(1, 0, 2), label=lab.reshape(T1, N * F).T)


def teacher_forced_loss(outputs, pad=0):
    """Trainer.compute_loss (trainer.py:61-79): (mean cross-entropy over non-PAD labels, token accuracy, argmax predictions [N*F, T-1])."""
    emb, ptr, labels = outputs["embedding"].astype(np.float64), outputs["pointer"].astype(np.float64), outputs["label"]
    logits = np.matmul(emb, ptr.transpose(0, 2, 1))                            # [NF, L, T-1]
    m = logits.max(axis=1, keepdims=True)
    logp = logits - m - np.log(np.exp(logits - m).sum(axis=1, keepdims=True))
    valid = labels != pad
    picked = np.take_along_axis(logp, labels[:, None, :], axis=1)[:, 0, :]
    loss = float(-(picked * valid).sum() / valid.sum())
    pred = logits.argmax(axis=1)
    acc = float((valid & (pred == labels)).sum() / (valid.sum() + 1e-10))
    return loss, acc, pred
