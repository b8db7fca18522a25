"""faceformer_b200 -- B200-native (sm_100a) greedy pointer-decode path of FaceFormer.

Only what the hot path needs:
  csrc/        hand-written CUDA kernels + the C ABI (include/ffb200.h)
  lib.py       ctypes binding of libffb200.so (no CPU fallback)
  engine.py    one handle per device: weights, encode, greedy decode, parity hooks
  models.py    SurfaceFormer_Parallel_B200 / SurfaceFormer_B200: the reference's model boundary
  sharding.py  wireframe-level data parallelism across ranks (NCCL broadcast / all-gather)
  synth.py     seeded synthetic wireframes and weights (numpy only)
  config.py    values of the reference's cfg.model for its shipped configs
"""
from .config import MODE_PARALLEL, MODE_SEQ2SEQ, OURS, OURS_PERSPECTIVE, SEQ2SEQ, TINY, ModelConfig  # noqa: F401

__all__ = ["ModelConfig", "MODE_PARALLEL", "MODE_SEQ2SEQ", "OURS", "OURS_PERSPECTIVE", "SEQ2SEQ", "TINY"]
