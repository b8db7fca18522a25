"""One launch of the CTA-pair GEMM (and of the single-CTA kernel) for ncu.  python profiles/probe_gemm2.py [variant]"""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa
from faceformer_b200.config import MODE_PARALLEL, OURS
from faceformer_b200.engine import Engine
from faceformer_b200.lib import FFB_OPT_GEMM_VARIANT
e = Engine(OURS, MODE_PARALLEL, 0)
for var in ([int(sys.argv[1])] if len(sys.argv) > 1 else [1, 3]):
    e.set_option(FFB_OPT_GEMM_VARIANT, var)
    ms = C.c_float()
    e._check(e._lib.ffb_bench_linear_tc(e._h, 131072, 512, 512, 1 | 32, 2, C.byref(ms), e._stream()))
    print(var, ms.value)
e.close()
