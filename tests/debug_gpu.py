"""Ad-hoc GPU diagnostics (not a test): eager determinism, graph-vs-eager, bf16-vs-f32 per layer."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import conftest  # noqa
import numpy as np, torch
from oracle import basi_oracle as O
from test_gpu_net import _setup, _engine, _rel

variant, nseg, S, F, B, classes = "2AddClass", 1, 64, 8, 2, 21
params, img, clicks, data, lab, cls, sigma = _setup(variant, nseg, S, F, B, classes)
loss = dict(kind="bce", pos_weight=3.0, class_weight=0.2)
e1 = _engine(variant, nseg, S, F, B, classes, "f32", loss)
e1b = _engine(variant, nseg, S, F, B, classes, "f32", loss)
e2 = _engine(variant, nseg, S, F, B, classes, "f32", loss)
for e in (e1, e1b, e2):
    e.set_params(params); e.feed(data, lab, cls, 1e-2)
e2.capture(train=True)
for step in range(3):
    e1.step_device(); e1b.step_device(); e2.replay()
    torch.cuda.synchronize()
    p1, p1b, p2 = e1.get_params(), e1b.get_params(), e2.get_params()
    g1, g1b, g2 = e1.get_grads(), e1b.get_grads(), e2.get_grads()
    w = max((_rel(p1[n], p1b[n]), n) for n in p1); w2 = max((_rel(p1[n], p2[n]), n) for n in p1)
    wg = max((_rel(g1[n], g1b[n]), n) for n in g1); wg2 = max((_rel(g1[n], g2[n]), n) for n in g1)
    print("step", step, "eager/eager params", w, "grads", wg)
    print("step", step, "eager/graph params", w2, "grads", wg2, "losses", e1.losses(), e2.losses())

# bf16 vs f32 per layer
for prec in ("bf16",):
    ef = _engine(variant, nseg, S, 16, B, classes, "f32", loss)
    eb = _engine(variant, nseg, S, 16, B, classes, prec, loss)
    params, img, clicks, data, lab, cls, sigma = _setup(variant, nseg, S, 16, B, classes)
    for e in (ef, eb):
        e.set_params(params); e.feed(data, lab, cls, 1e-2); e.step_device()
    torch.cuda.synchronize()
    names = ["conv1_1_3x3_s2_bn_relu", "conv1_3_3x3_bn", "pool1_3x3_s2", "conv2_1/relu", "conv2_3/relu", "conv3_1/relu", "conv3_4/relu",
             "conv4_1/relu", "conv4_6/relu", "conv4_12/relu", "conv4_23/relu", "conv5_1/relu", "conv5_3/relu", "conv5_4_bn", "conv6_n", "class_attention_fc"]
    for n in names:
        a = ef._acts[ef.net.layers[n].index]; b = eb._acts[eb.net.layers[n].index]
        x, y = a.t.float().cpu().numpy().astype(np.float64), b.t.float().cpu().numpy().astype(np.float64)
        print("%-28s rel-l2 %.3e  max-rel %.3e" % (n, np.linalg.norm(x - y) / np.linalg.norm(x), _rel(y, x)))
    gf, gb = ef.get_grads(), eb.get_grads()
    a = np.concatenate([gf[n].reshape(-1) for n in gf]); b = np.concatenate([gb[n].reshape(-1) for n in gf])
    print("grad cosine", float(a @ b / np.linalg.norm(a) / np.linalg.norm(b)), "losses", ef.losses(), eb.losses())
