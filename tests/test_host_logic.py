"""Host-side logic that needs no GPU: weight packing, synthetic generators, the model classes' state_dict
layout and the faceformer.models overlay (drop-in mechanism)."""
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

from faceformer_b200 import synth
from faceformer_b200.config import MODE_PARALLEL, MODE_SEQ2SEQ, OURS, SEQ2SEQ, TINY
from faceformer_b200.engine import pack_state_dict
from faceformer_b200.lib import FFBError

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_pack_state_dict_is_strict():
    sd = synth.synth_state_dict(TINY, MODE_PARALLEL, 0)
    blob = pack_state_dict(sd, TINY, MODE_PARALLEL)
    n_float = sum(v.size for k, v in sd.items() if v.dtype == np.float32)
    assert blob.dtype == np.float32 and blob.size == n_float
    assert np.array_equal(blob[:sd["val_enc.embedding_token.weight"].size], sd["val_enc.embedding_token.weight"].ravel())
    assert np.array_equal(blob[-TINY.num_model:], sd["project.bias"])
    # Lightning checkpoints prefix every key with "model." (trainer.py:20)
    assert np.array_equal(pack_state_dict({"model." + k: v for k, v in sd.items()}, TINY, MODE_PARALLEL), blob)
    bad = dict(sd); bad.pop("project.bias")
    with pytest.raises(FFBError, match="missing"):
        pack_state_dict(bad, TINY, MODE_PARALLEL)
    bad = dict(sd); bad["extra.weight"] = np.zeros(3, np.float32)
    with pytest.raises(FFBError, match="unexpected"):
        pack_state_dict(bad, TINY, MODE_PARALLEL)
    bad = dict(sd); bad["project.weight"] = np.zeros((3, 3), np.float32)
    with pytest.raises(FFBError, match="shape"):
        pack_state_dict(bad, TINY, MODE_PARALLEL)


def test_synth_is_deterministic_and_well_formed():
    a = synth.synth_batch(OURS, MODE_PARALLEL, 4, seed=5)
    b = synth.synth_batch(OURS, MODE_PARALLEL, 4, seed=5)
    for k in a:
        assert np.array_equal(a[k], b[k])
    assert a["input"].shape == (4, 216, 50, 2) and a["input"].dtype == np.float32
    assert a["input_mask"].dtype == np.bool_ and a["label"].shape == (4, 216, 37)
    for i, n in enumerate(a["num_input"]):
        assert 24 <= n <= 216
        assert not a["input_mask"][i, :n].any() and a["input_mask"][i, n:].all()
        assert np.all(a["input"][i, n:] == 0) and np.abs(a["input"][i, :n]).max() <= 1.0
    s = synth.synth_batch(SEQ2SEQ, MODE_SEQ2SEQ, 2, seed=1, num_edges=np.array([64, 3]))
    assert s["label"].shape == (2, 259)
    p = synth.polygon_batch(TINY, 3, seed=2)
    assert p["label"].shape == (3, 28, 10) and (p["label"][:, :, 0] > 0).all()


@pytest.mark.parametrize("mode,cfg", [(MODE_PARALLEL, TINY), (MODE_SEQ2SEQ, TINY)])
def test_b200_model_classes_have_the_reference_state_dict_layout(mode, cfg):
    from faceformer_b200.models import SurfaceFormer_B200, SurfaceFormer_Parallel_B200
    cls = SurfaceFormer_Parallel_B200 if mode == MODE_PARALLEL else SurfaceFormer_B200
    m = cls(**cfg.model_kwargs(mode), max_num_faces=42)
    sd = m.state_dict()
    want = synth.state_dict_names(cfg, mode)
    assert list(sd.keys()) == [n for n, _, _ in want]
    for n, shape, dt in want:
        assert tuple(sd[n].shape) == tuple(shape)
        assert sd[n].dtype == (torch.int64 if dt == "i8" else torch.float32)
    m.load_state_dict({k: torch.from_numpy(v) for k, v in synth.synth_state_dict(cfg, mode, 1).items()}, strict=True)
    # forward_train = the teacher-forced forward pass (row f4); like eval it has no CPU path
    with pytest.raises(FFBError, match="no CPU path"):
        m.train()({"input": torch.zeros(1, cfg.num_lines, 50, 2), "input_mask": torch.zeros(1, cfg.num_lines, dtype=torch.bool)})
    with pytest.raises(NotImplementedError):
        m.forward_train({}, scheduled_sampling_ratio=0.5)
    with pytest.raises(FFBError, match="no CPU path"):
        m.eval()({"input": torch.zeros(1, cfg.num_lines, 50, 2), "input_mask": torch.zeros(1, cfg.num_lines, dtype=torch.bool)})


@pytest.mark.skipif(not os.path.isdir("/root/reference/faceformer"), reason="reference tree not present on this box")
def test_overlay_shadows_faceformer_models_only():
    """main.py's `from faceformer.models import *` + str_to_class (main.py:9,13-14) must see the B200 classes while
    faceformer.embedding etc. still come from the reference (SURVEY.md 8b drop-in mechanism)."""
    code = ("import faceformer.models as m, faceformer.embedding as e, faceformer.transformer as t;"
            "from faceformer.models import *;"
            "print(m.__file__); print(e.__file__); print(SurfaceFormer_Parallel_B200.__module__, SurfaceFormer_Parallel.__module__)")
    env = dict(os.environ, PYTHONPATH=ROOT + os.pathsep + "/root/reference")
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=env, cwd="/tmp")
    assert out.returncode == 0, out.stderr
    l = out.stdout.strip().splitlines()
    assert l[0].startswith(ROOT) and l[1].startswith("/root/reference")
    assert l[2].split() == ["faceformer_b200.models", "faceformer.models.model_para"]


@pytest.mark.skipif(not os.path.isdir("/root/reference/faceformer"), reason="reference tree not present on this box")
def test_reference_state_dict_names_match_ours():
    sys.path.insert(0, "/root/reference")
    try:
        from faceformer.models.model_para import SurfaceFormer_Parallel
        ref = SurfaceFormer_Parallel(**TINY.model_kwargs(MODE_PARALLEL))
        want = [(k, tuple(v.shape)) for k, v in ref.state_dict().items()]
        assert want == [(n, tuple(s)) for n, s, _ in synth.state_dict_names(TINY, MODE_PARALLEL)]
    finally:
        sys.path.remove("/root/reference")
