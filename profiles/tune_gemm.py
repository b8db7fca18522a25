"""Microbenchmark of the tensor-core GEMM alone (ffb_bench_linear_tc).  Run on a GPU box:
    python profiles/tune_gemm.py
"""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: F401,E402  (creates the CUDA context / stream plumbing)

from faceformer_b200.config import MODE_PARALLEL, OURS  # noqa: E402
from faceformer_b200.engine import Engine  # noqa: E402

from faceformer_b200.lib import FFB_OPT_GEMM_VARIANT, FFB_OPT_TMA_EPILOGUE, FFB_OPT_TC_FORMAT  # noqa: E402

e = Engine(OURS, MODE_PARALLEL, 0)
RND = 32 if "--random" in sys.argv else 0        # pseudo-random operand bits instead of zeros (realistic power draw)
FLAGS = {"nostore": 16 | RND, "plain": 0 | RND, "bias": 1 | RND, "bias+res": 3 | RND, "bias+relu+split": 13 | RND,
         "dry bias+res": 67 | RND, "dry split": 77 | RND}     # dry = output formatted and staged, TMA stores not issued
for fmt, stag, var in ((2, 1, 0), (2, 1, 1), (2, 1, 3), (2, 0, 0), (3, 0, 0)):
  e.set_option(FFB_OPT_TC_FORMAT, fmt)
  e.set_option(FFB_OPT_TMA_EPILOGUE, stag)
  e.set_option(FFB_OPT_GEMM_VARIANT, var)
  print(f"--- operand format {fmt} ({'fp16x2, 3 MMA passes' if fmt == 2 else 'bf16x3, 6 MMA passes'}), TMA epilogue {stag}, variant {var}, {'random' if RND else 'zero'} operands")
  print(f"{'M':>8} {'N':>6} {'K':>6} " + " ".join(f"{k:>16}" for k in FLAGS))
  for M, N, K in [(131072, 512, 512), (131072, 1536, 512), (131072, 1024, 512), (131072, 512, 1024), (32768, 512, 512), (4096, 512, 512)]:
      row = []
      for name, fl in FLAGS.items():
          ms = C.c_float()
          e._check(e._lib.ffb_bench_linear_tc(e._h, M, N, K, fl, 10, C.byref(ms), e._stream()))
          row.append(f"{ms.value:7.3f}ms {2.0 * M * N * K / ms.value / 1e9:6.1f}TF")
      print(f"{M:>8} {N:>6} {K:>6} " + " ".join(f"{r:>16}" for r in row))
e.close()
