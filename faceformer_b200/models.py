"""Drop-in model classes for the reference's model boundary.

``SurfaceFormer_Parallel_B200`` / ``SurfaceFormer_B200`` take the reference constructors'
kwargs (model_para.py:14-19, model.py:14-18; built as ``model_class(**cfg.model)``,
trainer.py:20), hold parameters under the reference's ``state_dict`` names (strict
``load_state_dict`` of a reference checkpoint works) and implement
``forward(inputs: dict) -> dict`` for eval (model_para.py:243-259): the same dict comes back
with ``inputs['predict']`` (int64 [N,F,T] or [N,T]) on the input's device.

The modules below are parameter CONTAINERS only; no arithmetic of the path runs in PyTorch.
``forward`` hands device pointers to libffb200.so through ``Engine``.  ``forward_train``
(model_para.py:99-171) is the teacher-forced FORWARD pass only (loss / accuracy evaluation of a
batch, trainer.py:61-79): the library has no backward pass, so gradients do not flow.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from .config import MODE_PARALLEL, MODE_SEQ2SEQ, ModelConfig
from .engine import Engine
from .lib import FFBError


class _Attn(nn.Module):
    """Parameter layout of nn.MultiheadAttention (in_proj_weight/bias, out_proj.*)."""

    def __init__(self, E):
        super().__init__()
        self.in_proj_weight = nn.Parameter(torch.empty(3 * E, E))
        self.in_proj_bias = nn.Parameter(torch.zeros(3 * E))
        self.out_proj = nn.Linear(E, E)


class _EncLayer(nn.Module):
    def __init__(self, E, FF):
        super().__init__()
        self.self_attn = _Attn(E)
        self.linear1 = nn.Linear(E, FF)
        self.linear2 = nn.Linear(FF, E)
        self.norm1 = nn.LayerNorm(E)
        self.norm2 = nn.LayerNorm(E)


class _DecLayer(nn.Module):
    def __init__(self, E, FF):
        super().__init__()
        self.self_attn = _Attn(E)
        self.multihead_attn = _Attn(E)
        self.linear1 = nn.Linear(E, FF)
        self.linear2 = nn.Linear(FF, E)
        self.norm1 = nn.LayerNorm(E)
        self.norm2 = nn.LayerNorm(E)
        self.norm3 = nn.LayerNorm(E)


class _Stack(nn.Module):
    def __init__(self, layers, E):
        super().__init__()
        self.layers = nn.ModuleList(layers)
        self.norm = nn.LayerNorm(E)


class _ValEnc(nn.Module):
    def __init__(self, in_dim, E, num_token):
        super().__init__()
        self.embedding_token = nn.Embedding(num_token, E)
        self.embedding_value = nn.Sequential(nn.Linear(in_dim, E), nn.ReLU(), nn.Linear(E, E))


class _PosEnc(nn.Module):
    def __init__(self, E, max_len):
        super().__init__()
        self.register_buffer("position", torch.arange(0, max_len, dtype=torch.long).unsqueeze(0))
        self.pos_embed = nn.Embedding(max_len, E)


class _SurfaceFormerB200Base(nn.Module):
    MODE = MODE_PARALLEL

    def __init__(self, num_model=512, num_head=8, num_feedforward=2048, num_encoder_layers=6,
                 num_decoder_layers=6, dropout=0.1, activation="relu", normalize_before=True,
                 num_points_per_line=50, num_lines=64, point_dim=2, seq_length=10, token=None,
                 teacher_forcing_ratio=0, **kwargs):
        super().__init__()
        if activation != "relu" or not normalize_before:
            raise FFBError("only the reference's shipped configuration is supported: pre-norm + ReLU "
                           "(model_para.py:16; neither key is in faceformer/config.py)")
        num_token = int(token.len) if token is not None else 4
        kw = dict(num_model=num_model, num_head=num_head, num_feedforward=num_feedforward,
                  num_encoder_layers=num_encoder_layers, num_decoder_layers=num_decoder_layers,
                  num_points_per_line=num_points_per_line, num_lines=num_lines, point_dim=point_dim,
                  num_token=num_token, dropout=dropout)
        kw["max_face_length" if self.MODE == MODE_PARALLEL else "label_seq_length"] = seq_length
        self.cfg = ModelConfig(**kw)
        self.num_model = num_model
        self.token = token
        self.num_token = num_token
        E, FF = num_model, num_feedforward
        self.val_enc = _ValEnc(self.cfg.in_dim, E, num_token)
        self.pos_enc = _PosEnc(E, num_lines + num_token)
        self.query_pos_enc = _PosEnc(E, seq_length)
        self.encoder = _Stack([_EncLayer(E, FF) for _ in range(num_encoder_layers)], E)
        self.decoder = _Stack([_DecLayer(E, FF) for _ in range(num_decoder_layers)], E)
        self.project = nn.Linear(E, E)
        for _, p in self.named_parameters():               # model_para.py:50-53
            if p.dim() > 1:
                nn.init.xavier_uniform_(p)
        self._engines = {}
        self._weights_version = 0
        self.last_steps = None
        self.beam_width = 1           # > 1: beam search (parallel model; oracle/beam_oracle.py) -- 1 is the reference's greedy loop
        self.register_load_state_dict_post_hook(lambda module, incompatible: module.invalidate_weights())

    # -- weights -----------------------------------------------------------------------------
    def invalidate_weights(self):
        """Call after mutating parameters in place; load_state_dict does it automatically."""
        self._weights_version += 1

    def _apply(self, fn, *a, **k):
        out = super()._apply(fn, *a, **k)
        self._weights_version = getattr(self, "_weights_version", 0) + 1
        return out

    def engine(self, device_index: int) -> Engine:
        ent = self._engines.get(device_index)
        if ent is None:
            ent = [Engine(self.cfg, self.MODE, device_index), -1]
            self._engines[device_index] = ent
        if ent[1] != self._weights_version:
            ent[0].load_state_dict(self.state_dict())
            ent[1] = self._weights_version
        return ent[0]

    # -- forward -----------------------------------------------------------------------------
    def forward_train(self, inputs, scheduled_sampling_ratio=0):
        """Teacher-forced forward pass of forward_train (model_para.py:99-171 / model.py:98-157): sets the reference's outputs `embedding`
        ([N*F, L, E]: the encoder memory, replicated per anchor slot as model_para.py:116,164 do; [N, L, E] for seq2seq), `pointer`
        ([N*F, T-1, E] / [N, T-1, E]) and `label` ([N*F, T-1] / [N, T-1]), which Trainer.compute_loss consumes (trainer.py:61-79).  FORWARD ONLY: the tensors carry no autograd graph
        (there is no backward pass in libffb200), so this evaluates the training loss / token accuracy of a batch; optimisation steps
        need the reference classes.  Scheduled sampling (model_para.py:125-142) draws torch random numbers and is not offered."""
        if scheduled_sampling_ratio:
            raise NotImplementedError("scheduled sampling is not offered by the forward-only teacher-forced pass")
        coords = inputs["input"]
        if not (torch.is_tensor(coords) and coords.is_cuda):
            raise FFBError("faceformer_b200 has no CPU path: move the batch to a CUDA device")
        eng = self.engine(coords.device.index if coords.device.index is not None else torch.cuda.current_device())
        label, label_mask = inputs["label"], inputs["label_mask"]
        with torch.cuda.device(coords.device):
            if self.MODE == MODE_PARALLEL:
                num_input = inputs["num_input"]
                f = int(num_input.max().item())
                pointer, memory = eng.forward_train(coords.flatten(2), inputs["input_mask"], num_input, label, label_mask, want_embedding=True)
            else:
                pointer, memory = eng.forward_train(coords.flatten(2), inputs["input_mask"], None, label, label_mask, want_embedding=True)
            # memory [N, L, E] incl. the rows of padded edges (compute_loss's softmax runs over all L rows, trainer.py:64-69)
        if self.MODE == MODE_PARALLEL:
            inputs["embedding"] = memory.repeat_interleave(f, 0)                      # model_para.py:116,164
            inputs["label"] = label[:, :f, 1:].flatten(0, 1)                          # patch_target + flatten (model_para.py:84-86,166)
        else:
            inputs["embedding"] = memory                                              # model.py:154
            inputs["label"] = label[:, 1:]                                            # model.py:76-80,156
        inputs["pointer"] = pointer
        return inputs

    def forward_eval(self, inputs):
        coords = inputs["input"]
        if not (torch.is_tensor(coords) and coords.is_cuda):
            raise FFBError("faceformer_b200 has no CPU path: move the batch to a CUDA device "
                           "(Lightning does so before Trainer.forward, trainer.py:27-28)")
        eng = self.engine(coords.device.index if coords.device.index is not None else torch.cuda.current_device())
        num_input = inputs.get("num_input") if self.MODE == MODE_PARALLEL else None
        with torch.cuda.device(coords.device):
            if self.MODE == MODE_PARALLEL:
                from .lib import FFB_OPT_BEAM
                eng.set_option(FFB_OPT_BEAM, int(self.beam_width))
            predict, steps = eng.forward_eval(coords.flatten(2), inputs["input_mask"], num_input, want_steps=True)
            self.last_steps = steps
            if self.MODE == MODE_SEQ2SEQ:                   # model.py:216-217
                inputs["embedding"] = eng.get_memory()
                inputs["pointer"] = eng.get_last_pointer()
        inputs["predict"] = predict
        return inputs

    def forward(self, inputs):
        if self.training:
            return self.forward_train(inputs)
        return self.forward_eval(inputs)

    # -- the steps right before / after the path, on the same device ------
    def _current_engine(self) -> Engine:
        p = next(self.parameters())
        if not p.is_cuda:
            raise FFBError("faceformer_b200 has no CPU path: move the model to a CUDA device first")
        return self.engine(p.device.index if p.device.index is not None else torch.cuda.current_device())

    def featurize(self, wireframes):
        """Raw `edges` lists of the dataset JSON -> (input, input_mask, num_input) CUDA tensors, bit-identical to what
        ABCDataset_Parallel.__getitem__ + the default collate produce (datasets/data_para.py:8-25,59-68)."""
        return self._current_engine().featurize(wireframes, device=True)

    def parse_faces(self, predict, wireframes, tol: float = 2e-4, check_enclosed: bool = True):
        """inputs['predict'] + raw `edges` lists -> per wireframe the list of (face_type, loops) that parse_parallel_faces
        (trainer.py:196-206) followed by filter_faces_by_encloseness (post_processing.py:8-20) return for the predictions."""
        return self._current_engine().parse_faces(predict, wireframes, tol=tol, check_enclosed=check_enclosed)


    def face_accuracy(self, outputs, raw_datas, is_coedge: bool = True, tol: float = 2e-4):
        """Trainer.face_accuracy (trainer.py:210-300) on the dict forward() returned: per-sequence parsing on the GPU, set logic on the host."""
        from . import metrics
        return metrics.face_accuracy(self._current_engine(), outputs, raw_datas, is_coedge=is_coedge, tol=tol)


class SurfaceFormer_Parallel_B200(_SurfaceFormerB200Base):
    """Replaces SurfaceFormer_Parallel (model_para.py:12-259) on the eval path."""
    MODE = MODE_PARALLEL

    def __init__(self, *args, max_face_length=10, **kwargs):
        kwargs.pop("seq_length", None)
        super().__init__(*args, seq_length=max_face_length, **kwargs)
        self.max_face_length = max_face_length


class SurfaceFormer_B200(_SurfaceFormerB200Base):
    """Replaces SurfaceFormer (model.py:12-237) on the eval path."""
    MODE = MODE_SEQ2SEQ

    def __init__(self, *args, label_seq_length=2000, num_lines=1000, **kwargs):
        kwargs.pop("seq_length", None)
        super().__init__(*args, seq_length=label_seq_length, num_lines=num_lines, **kwargs)
        self.num_labels = label_seq_length
