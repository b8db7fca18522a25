"""input featurisation (datasets/data_para.py:8-25,59-68).  The oracle restatement is pinned to outputs of the
reference's own sample_points (tests/golden/featurize.npz); the CUDA kernel must reproduce them BIT-exactly."""
import json
import os

import numpy as np
import pytest

from oracle import featurize_oracle as fo

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "featurize.npz")


def cases():
    with np.load(GOLDEN) as z:
        meta = json.loads(str(z["meta"]))
        return {k: (c, z[f"{k}_input"], z[f"{k}_mask"]) for k, c in meta["cases"].items()}


@pytest.mark.parametrize("name", ["a", "b", "c"])
def test_oracle_matches_reference_functions(name):
    c, want, wmask = cases()[name]
    wfs = fo.synth_wireframes(c["n"], c["num_lines"], c["seed"])
    inp, mask, ni = fo.featurize(wfs, c["num_lines"])
    assert inp.dtype == np.float32 and np.array_equal(inp.view(np.uint32), want.view(np.uint32))
    assert np.array_equal(mask, wmask) and np.array_equal(ni, (~wmask).sum(1))


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["a", "b", "c"])
@pytest.mark.parametrize("device", [True, False])
def test_cuda_featurize_is_bit_exact(name, device):
    from faceformer_b200.config import MODE_PARALLEL, OURS
    from faceformer_b200.engine import Engine
    c, want, wmask = cases()[name]
    wfs = fo.synth_wireframes(c["n"], c["num_lines"], c["seed"])
    e = Engine(OURS.replace(num_lines=c["num_lines"]), MODE_PARALLEL, 0)
    inp, mask, ni = e.featurize(wfs, device=device)
    if device:
        inp, mask, ni = inp.cpu().numpy(), mask.cpu().numpy(), ni.cpu().numpy()
    assert inp.shape == want.shape and np.array_equal(inp.view(np.uint32), want.view(np.uint32))
    assert np.array_equal(mask, wmask) and np.array_equal(ni, (~wmask).sum(1))
    e.close()


@pytest.mark.gpu
def test_cuda_featurize_rejects_what_the_reference_rejects():
    from faceformer_b200.config import MODE_PARALLEL, TINY
    from faceformer_b200.engine import Engine
    from faceformer_b200.lib import FFBError
    e = Engine(TINY, MODE_PARALLEL, 0)
    seg = [[0.0, 0.0], [1.0, 1.0]]
    with pytest.raises(FFBError):
        e.featurize([[seg] * (TINY.num_lines + 1)])          # more edges than num_lines: IndexError in the reference
    e.close()


@pytest.mark.gpu
def test_featurize_feeds_the_decode_path():
    """featurize -> forward_eval on device tensors equals forward_eval on the oracle's featurised arrays."""
    import torch
    from faceformer_b200 import synth
    from faceformer_b200.config import MODE_PARALLEL, TINY
    from faceformer_b200.engine import Engine
    wfs = fo.synth_wireframes(4, TINY.num_lines, 9, lo=4)
    sd = synth.synth_state_dict(TINY, MODE_PARALLEL, 3, "diverse")
    e = Engine(TINY, MODE_PARALLEL, 0)
    e.load_state_dict(sd)
    inp, mask, ni = e.featurize(wfs)
    p1, s1 = e.forward_eval(inp.flatten(2), mask, ni)
    oi, om, on = fo.featurize(wfs, TINY.num_lines)
    p2, s2 = e.forward_eval(torch.from_numpy(oi).cuda().flatten(2), torch.from_numpy(om).cuda(), torch.from_numpy(on).cuda())
    assert s1 == s2 and torch.equal(p1, p2)
    e.close()
