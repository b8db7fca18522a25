"""Model/geometry configuration for the FaceFormer greedy pointer-decode path.

Mirrors the keys of the reference's ``cfg.model`` node
(/root/reference/faceformer/config.py:27-49) and the per-config YAML overrides
(/root/reference/configs/ours.yml:19-26, configs/seq2seq.yml:12-15,
configs/ours-perspective.yml:19-26).  Only values are mirrored; the fvcore
CfgNode mechanism is out of scope.
"""
from __future__ import annotations

from dataclasses import dataclass, asdict, replace
from types import SimpleNamespace

MODE_PARALLEL = 0   # SurfaceFormer_Parallel  (model_para.py:181-241)
MODE_SEQ2SEQ = 1    # SurfaceFormer           (model.py:169-219)


@dataclass(frozen=True)
class ModelConfig:
    num_model: int = 512            # E   (config.py:34)
    num_head: int = 8               # H   (config.py:35)
    num_feedforward: int = 1024     # FF  (config.py:36)
    num_encoder_layers: int = 6     # config.py:37
    num_decoder_layers: int = 6     # config.py:38
    num_points_per_line: int = 50   # config.py:28
    num_lines: int = 64             # config.py:29
    point_dim: int = 2              # config.py:30
    label_seq_length: int = 128     # seq2seq T (config.py:31)
    max_face_length: int = 34       # parallel T (config.py:33)
    num_token: int = 4              # token.len (config.py:47)
    dropout: float = 0.2            # identity in eval (config.py:39)

    @property
    def mem_len(self) -> int:
        """L = num_lines + 4 special-token rows (model_para.py:31)."""
        return self.num_lines + self.num_token

    @property
    def in_dim(self) -> int:
        return self.num_points_per_line * self.point_dim

    def seq_len(self, mode: int) -> int:
        """T: decoded sequence length incl. the start token."""
        return self.max_face_length if mode == MODE_PARALLEL else self.label_seq_length

    def token_namespace(self) -> SimpleNamespace:
        """The ``token`` object the reference constructors expect (config.py:40-48)."""
        return SimpleNamespace(PAD=0, SOS=1, SEP=2, EOS=3, DIR0=4, DIR1=5,
                               len=self.num_token, face_type_offset=1)

    def model_kwargs(self, mode: int) -> dict:
        """kwargs as ``Trainer`` splats them into the model (trainer.py:20)."""
        d = dict(num_model=self.num_model, num_head=self.num_head,
                 num_feedforward=self.num_feedforward,
                 num_encoder_layers=self.num_encoder_layers,
                 num_decoder_layers=self.num_decoder_layers,
                 dropout=self.dropout,
                 num_points_per_line=self.num_points_per_line,
                 num_lines=self.num_lines, point_dim=self.point_dim,
                 token=self.token_namespace())
        if mode == MODE_PARALLEL:
            d["max_face_length"] = self.max_face_length
        else:
            d["label_seq_length"] = self.label_seq_length
        return d

    def to_dict(self) -> dict:
        return asdict(self)

    def replace(self, **kw) -> "ModelConfig":
        return replace(self, **kw)


# The reference's shipped configurations (values only).
OURS = ModelConfig(num_lines=216, max_face_length=37)                  # configs/ours.yml:20-22
OURS_PERSPECTIVE = ModelConfig(num_lines=202, max_face_length=38)      # configs/ours-perspective.yml:20-22
SEQ2SEQ = ModelConfig(num_lines=110, label_seq_length=259)             # configs/seq2seq.yml:13-14
# Small geometry used by fast parity tests (same code paths, d_head stays 64).
TINY = ModelConfig(num_model=128, num_head=2, num_feedforward=256,
                   num_encoder_layers=2, num_decoder_layers=2,
                   num_lines=28, max_face_length=10, label_seq_length=24)
# E = 512 / H = 8 geometry (on the tcgen05 grid: E, FF multiples of 256) small enough to TRAIN on the CPU with the reference's own
# forward_train and to commit the checkpoint as an int8-grid fixture (oracle/train_fixture.py --cfg mid).
MID = ModelConfig(num_model=512, num_head=8, num_feedforward=256,
                  num_encoder_layers=1, num_decoder_layers=2,
                  num_lines=28, max_face_length=10, label_seq_length=24)
