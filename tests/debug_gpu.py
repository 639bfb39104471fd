"""Ad-hoc GPU diagnostics (not a test): bf16-vs-f32 per layer at the full benchmark size."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import conftest  # noqa
import numpy as np, torch
from oracle import basi_oracle as O
from test_gpu_net import _setup, _engine, _rel

variant, nseg, S, F, B, classes = "2AddClass", 1, 320, 32, 16, 21
loss = dict(kind="bce", pos_weight=3.0, class_weight=0.2)
for trained in (True, False):
    params, img, clicks, data, lab, cls, sigma = _setup(variant, nseg, S, F, B, classes)
    if not trained:
        params = O.init_params(O.param_specs(variant, classes, nseg, F), 1, trained_like=False)
    ef = _engine(variant, nseg, S, F, B, classes, "f32", loss)
    eb = _engine(variant, nseg, S, F, B, classes, "bf16", loss)
    for e in (ef, eb):
        e.set_params(params); e.feed(data, lab, cls, 1e-2); e.step_device()
    torch.cuda.synchronize()
    names = ["conv1_3_3x3_bn", "conv2_3/relu", "conv3_4/relu", "conv4_6/relu", "conv4_23/relu", "conv5_3/relu", "conv5_4_bn", "conv6_n", "class_attention_fc"]
    print("trained_like", trained)
    for n in names:
        a = ef._acts[ef.net.layers[n].index]; b = eb._acts[eb.net.layers[n].index]
        x, y = a.t.float().cpu().numpy().astype(np.float64), b.t.float().cpu().numpy().astype(np.float64)
        print("  %-24s rel-l2 %.3e  max-rel %.3e" % (n, np.linalg.norm(x - y) / np.linalg.norm(x), _rel(y, x)))
    gf, gb = ef.get_grads(), eb.get_grads()
    a = np.concatenate([gf[n].reshape(-1) for n in gf]); b = np.concatenate([gb[n].reshape(-1) for n in gf])
    pf = (ef.seg_logits.t.cpu().numpy() > 0); pb = (eb.seg_logits.t.cpu().numpy() > 0)
    iou = (pf & pb).sum() / max(1, (pf | pb).sum())
    print("  grad cosine %.4f rel-l2 %.3e | losses f32 %s bf16 %s | mask(sigmoid>0.5) IoU bf16-vs-f32 %.4f agree %.4f" % (
        float(a @ b / np.linalg.norm(a) / np.linalg.norm(b)), np.linalg.norm(a - b) / np.linalg.norm(a), ef.losses(), eb.losses(), iou, (pf == pb).mean()))
    del ef, eb
    torch.cuda.empty_cache()
