"""Per-stage error budget of the pointer logits, reference side (TEST INFRASTRUCTURE; build container only).

For a few golden cases: (1) the reference's encoder memory evaluated in float64 (valid rows only, rounded to fp32) -- feeding it
to the CUDA path through ffb_set_memory isolates the decoder + pointer-head contribution of OUR logit error; (2) the reference's
own fp32-vs-float64 logit distance split by stage, obtained by running exactly one stage of the unmodified reference in fp32
and the others in float64 (encoder / decoder stack / project + pointer dot).

    python oracle/make_golden_noise.py      # writes tests/golden/noise_budget.npz
"""
from __future__ import annotations

import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference")
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))

from faceformer_b200.config import MODE_PARALLEL  # noqa: E402
from make_golden import GOLDEN, build_reference  # noqa: E402
from util import load_case  # noqa: E402

CASES = ["ours_parallel_small", "seq2seq_single64", "perspective_small"]


def staged_logits(batch, mode, prefix, m_enc, m_dec, m_head):
    """Loop body of forward_eval (model_para.py:191-227) with each stage run by its own copy of the reference model."""
    def tb(m):
        d = next(m.parameters()).dtype
        return {k: (torch.from_numpy(v).to(d) if v.dtype == np.float32 else torch.from_numpy(v)) for k, v in batch.items()}
    with torch.no_grad():
        b = tb(m_enc)
        input_mask = m_enc.process_masks(b["input_mask"])
        val, pos, _ = m_enc.get_embeddings(b["input"], b["label"])
        source, pos = m_enc.patch_source(val, pos)
        memory = m_enc.encoder(source, src_key_padding_mask=input_mask, pos=pos)
        mem_out = memory.transpose(0, 1).double().numpy()
        dd = next(m_dec.parameters()).dtype
        memory = memory.to(dd)
        b2 = tb(m_dec)
        _, pos2, qpos2 = m_dec.get_embeddings(b2["input"], b2["label"])
        _, pos2 = m_dec.patch_source(val.to(dd), pos2)
        qpos2 = qpos2.transpose(0, 1)
        if mode == MODE_PARALLEL:
            F = int(max(b["num_input"]))
            memory = memory.repeat_interleave(F, 1)
            input_mask = input_mask.repeat_interleave(F, 0)
        pre = torch.from_numpy(prefix)
        tgt = torch.gather(memory, 0, pre.unsqueeze(-1).repeat(1, 1, m_dec.num_model))
        hid = m_dec.decoder(tgt, memory, memory_key_padding_mask=input_mask, pos=pos2, query_pos=qpos2[:pre.size(0)])
        dh = next(m_head.parameters()).dtype
        ptr = m_head.project(hid.to(dh))
        logit = torch.bmm(memory.to(dh).transpose(0, 1), ptr.permute(1, 2, 0)[..., -1:])[..., 0]
    return logit.double().numpy(), input_mask.numpy(), mem_out


def main():
    torch.set_num_threads(os.cpu_count())
    out, budget = {}, {}
    for name in CASES:
        g = load_case(name)
        cfg, mode, sd, batch = g["cfg"], g["mode"], g["sd"], g["batch"]
        m32, m64 = build_reference(cfg, mode, sd), build_reference(cfg, mode, sd).double()
        T = cfg.seq_len(mode)
        prefix = np.ascontiguousarray(g["predict"].reshape(-1, T)[:, :g["steps"]].T)
        ref, mask, mem64 = staged_logits(batch, mode, prefix, m64, m64, m64)

        def d(x):
            return float(np.abs(x - ref)[~mask].max())
        b = {"max_logit": float(np.abs(ref[~mask]).max()),
             "all_fp32": d(staged_logits(batch, mode, prefix, m32, m32, m32)[0]),
             "encoder_fp32": d(staged_logits(batch, mode, prefix, m32, m64, m64)[0]),
             "decoder_fp32": d(staged_logits(batch, mode, prefix, m64, m32, m64)[0]),
             "head_fp32": d(staged_logits(batch, mode, prefix, m64, m64, m32)[0])}
        budget[name] = b
        valid = np.concatenate([np.ones((mem64.shape[0], cfg.num_token), bool), ~batch["input_mask"]], 1)
        mem = np.where(valid[..., None], mem64, 0.0).astype(np.float32)        # [N, L, E]; padded rows zeroed (never read)
        out[name + "::memory64"] = mem
        print(name, json.dumps(b), flush=True)
    np.savez_compressed(os.path.join(GOLDEN, "noise_budget.npz"), meta=json.dumps(budget), **out)


if __name__ == "__main__":
    main()
