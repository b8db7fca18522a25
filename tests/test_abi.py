"""The C-ABI library loads and exports every symbol include/ffb200.h declares.  No compute calls."""
import ctypes as C
import os
import re

import pytest
import torch

from faceformer_b200 import lib as L
from faceformer_b200.config import MODE_PARALLEL, MODE_SEQ2SEQ, OURS, SEQ2SEQ, TINY
from faceformer_b200.synth import state_dict_names

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _cfg(cfg, mode_, **over):
    d = dict(abi_version=L.FFB_ABI_VERSION, mode=mode_, num_model=cfg.num_model, num_head=cfg.num_head,
             num_feedforward=cfg.num_feedforward, num_encoder_layers=cfg.num_encoder_layers,
             num_decoder_layers=cfg.num_decoder_layers, in_dim=cfg.in_dim, num_lines=cfg.num_lines,
             num_token=cfg.num_token, seq_len=cfg.seq_len(mode_), device=0)
    d.update(over)
    return L.ffb_config(**d)


def test_library_is_built_in_tree():
    L.build()
    assert os.path.exists(L.LIB_PATH) and os.path.dirname(L.LIB_PATH).endswith("faceformer_b200")


def test_header_symbols_are_all_exported_and_bound():
    hdr = open(os.path.join(ROOT, "include", "ffb200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(ffb_[a-z0-9_]+)\s*\(", hdr))
    assert declared, "no declarations found"
    lib = L.load()
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in ffb200.h but not exported"
    assert declared == set(L.SIGNATURES), (declared ^ set(L.SIGNATURES))


@pytest.mark.parametrize("cfg,mode", [(OURS, MODE_PARALLEL), (SEQ2SEQ, MODE_SEQ2SEQ), (TINY, MODE_PARALLEL)])
def test_weight_count_matches_reference_state_dict(cfg, mode):
    want = sum(int(torch.Size(shape).numel()) for _, shape, dt in state_dict_names(cfg, mode) if dt == "f4")
    c = _cfg(cfg, mode)
    assert L.load().ffb_weight_count(C.byref(c)) == want
    if cfg is OURS:
        assert want == 32_256_000


@pytest.mark.parametrize("over", [dict(num_model=500), dict(num_head=4), dict(in_dim=99), dict(abi_version=99),
                                  dict(mode=7), dict(seq_len=1), dict(num_token=3)])
def test_invalid_config_is_rejected(over):
    c = _cfg(OURS, MODE_PARALLEL, **over)
    lib = L.load()
    assert lib.ffb_weight_count(C.byref(c)) == 0
    h = C.c_void_p()
    assert lib.ffb_create(C.byref(c), C.byref(h)) == -1 and not h.value
    assert b"invalid config" in lib.ffb_last_error(None)


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_create_fails_loudly_without_gpu():
    """No CPU fallback: without a CUDA device the product refuses to construct."""
    c = _cfg(TINY, MODE_PARALLEL)
    lib = L.load()
    h = C.c_void_p()
    st = lib.ffb_create(C.byref(c), C.byref(h))
    assert st == -2 and not h.value
    assert b"no CPU fallback" in lib.ffb_last_error(None)
    from faceformer_b200.engine import Engine
    with pytest.raises(L.FFBError):
        Engine(TINY, MODE_PARALLEL)
