"""Golden fixtures (tests/golden/*.npz, made by tests/golden/make_golden.py from the float64 oracle).

CPU: the oracle still reproduces them.  GPU: the CUDA path (fp32 mode) reproduces them through the engine."""
import os
import sys

import numpy as np
import pytest
import torch

from oracle import basi_oracle as O

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
import make_golden as G  # noqa: E402


def _load(name):
    return dict(np.load(os.path.join(HERE, "golden", "golden_%s.npz" % name)))


@pytest.mark.parametrize("name", sorted(G.CASES))
def test_oracle_reproduces_golden(name):
    g, now = _load(name), G.build(name)
    assert np.array_equal(g["images"], now["images"]) and np.array_equal(g["clicks"], now["clicks"])
    assert np.array_equal(g["click_map"].view(np.uint32), now["click_map"].view(np.uint32))     # bit exact
    assert abs(float(g["loss"]) - float(now["loss"])) < 1e-10
    assert np.allclose(g["seg_logits"], now["seg_logits"], rtol=0, atol=1e-6)
    for k in G.GRAD_KEYS:
        assert np.allclose(g["grad:" + k], now["grad:" + k], rtol=1e-5, atol=1e-9), k


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(G.CASES))
def test_cuda_path_reproduces_golden(name):
    from basi_b200.BAISPSPNet import PSPNet, Placeholder
    from basi_b200.engine import Engine
    variant, nseg, S, F, B, classes, pw, cw = G.CASES[name]
    g = _load(name)
    params = G.inputs(name)[4]
    net = PSPNet({'data': Placeholder((None, S, S, 4))}, num_classes=classes, num_segment=nseg, is_training=True,
                 last_pool_size=S // 8, filter_number=F, variant=variant)
    eng = Engine(net, B, "f32", True, dict(kind="bce" if nseg == 1 else "softmax", pos_weight=pw, class_weight=cw))
    eng.set_params(params)
    eng.enable_click_input(30)
    eng.feed_clicks(g["images"], g["clicks"])
    eng.feed(None, g["label_seg"], g["label_cls"], 5e-3)
    eng.step_device()
    torch.cuda.synchronize()
    assert np.array_equal(eng.input.t.cpu().numpy()[..., 3].view(np.uint32), g["click_map"].view(np.uint32))
    total, lseg, lcls = eng.losses()
    assert abs(total - float(g["loss"])) < 1e-4 * max(1.0, abs(float(g["loss"])))
    logits = eng.seg_logits.t.cpu().numpy()
    assert np.max(np.abs(logits - g["seg_logits"])) < 1e-4 * np.max(np.abs(g["seg_logits"]))
    grads = eng.get_grads()
    for k in G.GRAD_KEYS:
        ref = g["grad:" + k].astype(np.float64)
        err = np.linalg.norm(grads[k].astype(np.float64).reshape(-1) - ref.reshape(-1)) / max(np.linalg.norm(ref), 1e-30)
        assert err < 1e-4 + 10 * float(g["floor:" + k]), (k, err, float(g["floor:" + k]))
