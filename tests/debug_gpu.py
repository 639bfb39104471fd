"""Ad-hoc GPU diagnostics (not a test)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import conftest  # noqa
import numpy as np, torch
from oracle import basi_oracle as O
from test_gpu_net import _setup, _engine, _rel, CASES

for case in (CASES[4], ("1NoClass", 1, 320, 8, 1, 21, 3.0, 0.0), ("2AddClass", 1, 320, 8, 2, 21, 5.0, 0.1)):
    variant, nseg, S, F, B, classes, pw, cw = case
    params, img, clicks, data, lab, cls, sigma = _setup(variant, nseg, S, F, B, classes)
    eng = _engine(variant, nseg, S, F, B, classes, "f32", dict(kind="bce", pos_weight=pw, class_weight=cw))
    eng.set_params(params); eng.feed(data, lab, cls, 5e-3); eng.step_device(); torch.cuda.synchronize()
    ref = O.train_step(params, data, lab, cls, variant, nseg, S // 8, pw, cw, 5e-3, torch.float64)
    r32 = O.train_step(params, data, lab, cls, variant, nseg, S // 8, pw, cw, 5e-3, torch.float32)
    g = eng.get_grads()
    rows = []
    for n in g:
        if np.max(np.abs(ref["grads"][n])) <= 1e-12: continue
        rows.append((_rel(g[n], ref["grads"][n]), _rel(r32["grads"][n], ref["grads"][n]), float(np.max(np.abs(ref["grads"][n]))), n))
    rows.sort(reverse=True)
    print(case, "logits", _rel(eng.seg_logits.t.cpu().numpy(), ref["seg_logits"]), "floor", _rel(r32["seg_logits"], ref["seg_logits"]))
    for r in rows[:12]:
        print("   err %.2e floor %.2e |g|max %.2e %s" % r)
    bad = [r for r in rows if r[0] > 1e-4 + 10 * r[1]]
    print("   failing:", len(bad), "of", len(rows))
