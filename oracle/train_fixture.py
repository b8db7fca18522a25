"""Train a small NON-DEGENERATE FaceFormer checkpoint with the reference's own code.

TEST INFRASTRUCTURE (runs only in the build container; needs /root/reference).

Random-init weights make greedy decode collapse to one repeated token
(SURVEY.md section 7 "Degenerate synthetic weights"), which would make
end-to-end token parity nearly vacuous.  This script runs the UNMODIFIED
reference model's teacher-forced ``forward_train`` (model_para.py:99-171) with
the loss of ``Trainer.compute_loss`` (trainer.py:60-80, restated because
pytorch_lightning is not installed) on synthetic closed-polygon wireframes and
writes the resulting ``state_dict`` as a small fp32 ``.npz`` under
``tests/golden/``.  Labels are built exactly as
``ABCDataset_Parallel.__getitem__`` does (data_para.py:70-95; restated in
``faceformer_b200.synth.polygon_sample``).

    python oracle/train_fixture.py --steps 1500 --out tests/golden/tiny_trained_parallel.npz
"""
from __future__ import annotations

import argparse
import os
import sys
import time

import numpy as np
import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference")

from faceformer_b200.config import MID, MODE_PARALLEL, TINY  # noqa: E402
from faceformer_b200.synth import load_state_dict_npz, polygon_sample, quantize_state_dict  # noqa: E402


def collate(samples):
    out = {}
    for k in samples[0]:
        out[k] = torch.from_numpy(np.stack([np.asarray(s[k]) for s in samples]))
    return out


def compute_loss(outputs):
    """trainer.py:60-80."""
    logits = torch.bmm(outputs["embedding"], outputs["pointer"].transpose(1, 2))
    labels = outputs["label"].detach().clone()
    loss = F.cross_entropy(logits, labels, ignore_index=0, reduction="sum")
    valid = labels != 0
    acc = float((valid * (logits.argmax(1) == labels)).sum()) / float(valid.sum() + 1e-10)
    return loss / valid.sum(), acc


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=1500)
    ap.add_argument("--batch", type=int, default=8)
    ap.add_argument("--lr", type=float, default=1e-3)
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--max-seconds", type=float, default=1500)
    ap.add_argument("--out", default=os.path.join(ROOT, "tests/golden/tiny_trained_parallel.npz"))
    ap.add_argument("--cfg", default="tiny", choices=["tiny", "mid"],
                    help="tiny: E=128 (fp32 checkpoint); mid: E=512/H=8 on the tcgen05 grid, saved on an int8 grid (--quantize 8)")
    ap.add_argument("--quantize", type=int, default=0, help="snap >= 2-d weights to a per-tensor symmetric N-bit grid before saving")
    ap.add_argument("--warmup-steps", type=int, default=0, help="linear learning-rate warm-up")
    args = ap.parse_args()

    from faceformer.models import SurfaceFormer_Parallel
    cfg = MID if args.cfg == "mid" else TINY
    torch.manual_seed(args.seed)
    torch.set_num_threads(os.cpu_count())
    rng = np.random.default_rng(args.seed)
    kw = cfg.model_kwargs(MODE_PARALLEL); kw["dropout"] = 0.0
    model = SurfaceFormer_Parallel(**kw).train()
    opt = torch.optim.Adam(model.parameters(), lr=args.lr)
    t0 = time.time()
    for step in range(args.steps):
        batch = collate([polygon_sample(rng, cfg) for _ in range(args.batch)])
        out = model(batch)
        loss, acc = compute_loss(out)
        if args.warmup_steps:
            for g in opt.param_groups:
                g["lr"] = args.lr * min(1.0, (step + 1) / args.warmup_steps)
        opt.zero_grad(); loss.backward()
        torch.nn.utils.clip_grad_norm_(model.parameters(), 1.0)
        opt.step()
        if step % 50 == 0 or step == args.steps - 1:
            print(f"step {step} loss {loss.item():.4f} acc {acc:.3f} t {time.time()-t0:.0f}s", flush=True)
        if time.time() - t0 > args.max_seconds:
            break
    model.eval()
    sd = {k: v.detach().cpu().numpy() for k, v in model.state_dict().items()}
    if args.quantize:
        np.savez_compressed(args.out, **quantize_state_dict(sd, args.quantize))
        sd = load_state_dict_npz(args.out)          # evaluate what the fixture actually holds
        model.load_state_dict({k: torch.from_numpy(np.asarray(v)) for k, v in sd.items()}, strict=True)
    else:
        np.savez_compressed(args.out, **sd)
    # quick look at greedy behaviour
    batch = collate([polygon_sample(rng, cfg) for _ in range(2)])
    with torch.no_grad():
        pred = model(batch)["predict"].numpy()
    print("predict[0][:8]:\n", pred[0][:8])
    print("label[0][:8]:\n", batch["label"][0][:8].numpy())
    print("saved", args.out, os.path.getsize(args.out) / 1e6, "MB")


if __name__ == "__main__":
    main()
