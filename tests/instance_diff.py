"""Diagnostic (run by hand on a GPU box): builds two engines on the same parameters / batch and lists the gradient
tensors that differ most between them, plus the same comparison for two steps of ONE engine.

  python tests/instance_diff.py [f16|bf16] [variant]
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "instance-segment-basi_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
from test_gpu_net import _engine, _setup  # noqa: E402

prec = sys.argv[1] if len(sys.argv) > 1 else "bf16"
variant = sys.argv[2] if len(sys.argv) > 2 else "2AddClass"
nseg, S, F, B, classes = 1, 320, 32, 4, 21
params, img, clicks, data, lab, cls, sigma = _setup(variant, nseg, S, F, B, classes)
cw = 0.0 if variant == "1NoClass" else 0.2


def run(eng):
    eng.set_params(params)
    eng.feed(data, lab, cls if cw else None, 5e-3)
    eng.step_device()
    torch.cuda.synchronize()
    acts = {}
    for name in PROBE:
        try:
            acts[name] = eng.fetch(name, grad=True).copy()
        except Exception as e:          # not materialised / no gradient buffer
            acts[name] = None
    ACTS.append(acts)
    fw = {}
    for name in FPROBE:
        try:
            fw[name] = eng.fetch(name).copy()
        except Exception:
            fw[name] = None
    FWD.append(fw)
    return eng.get_grads(), eng.seg_logits.t.float().cpu().numpy().copy()


PROBE = ["class_attention_fc", "class_attention_squeeze", "class_attention_conv", "class_attention_pool",
         "class_attention_multiply", "conv6_n", "conv5_4_bn", "conv5_4", "conv5_3_concat", "conv5_3/relu", "conv5_3",
         "conv5_3_pool1_interp", "conv5_3_pool1_conv_bn", "conv5_3_pool1_conv", "conv5_3_pool1", "conv5_3_pool6_conv",
         "conv5_3_1x1_increase", "conv5_3_3x3", "conv5_2/relu", "conv4_23/relu"]
ACTS = []
FPROBE = ["conv1_1_3x3_s2_bn", "conv1_2_3x3_bn", "conv1_3_3x3_bn", "pool1_3x3_s2", "conv2_1_1x1_reduce_bn", "conv2_1_3x3_bn",
          "conv2_1/relu", "conv2_3/relu", "conv3_1/relu", "conv3_4/relu", "conv4_1/relu", "conv4_12/relu", "conv4_23/relu",
          "conv5_1/relu", "conv5_3/relu", "conv5_3_pool1", "conv5_3_pool2", "conv5_3_pool3", "conv5_3_pool6",
          "conv5_3_pool1_conv_bn", "conv5_3_pool6_conv_bn", "conv5_3_pool1_interp", "conv5_3_pool6_interp", "conv5_4_bn", "conv6_n",
          "class_attention_multiply", "class_attention_pool", "class_attention_conv", "class_attention_fc"]
FWD = []


def fwd_report(tag, a, b):
    print("%s: forward activations:" % tag)
    for name in FPROBE:
        if a.get(name) is None or b.get(name) is None:
            print("   %-28s (not materialised)" % name)
            continue
        x, y = a[name].astype(np.float64), b[name].astype(np.float64)
        print("   %-28s max |diff| %.3e   differing elements %d of %d" % (name, np.abs(x - y).max(), int((x != y).sum()), x.size))


def probe_report(tag, a, b):
    print("%s: activation gradients (first differing from the loss backwards):" % tag)
    for name in PROBE:
        if a.get(name) is None or b.get(name) is None:
            print("   %-28s (not materialised)" % name)
            continue
        x, y = a[name].astype(np.float64), b[name].astype(np.float64)
        print("   %-28s rel-l2 %.3e   |g| %.3e" % (name, np.linalg.norm(x - y) / max(np.linalg.norm(x), 1e-30), np.linalg.norm(x)))


def report(tag, ga, gb):
    rows = []
    for n in ga:
        a, b = ga[n].astype(np.float64).ravel(), gb[n].astype(np.float64).ravel()
        rows.append((np.linalg.norm(a - b), np.linalg.norm(a), n))
    tot = np.sqrt(sum(r[0] ** 2 for r in rows)) / np.sqrt(sum(r[1] ** 2 for r in rows))
    print("%s: total gradient rel-l2 %.3e; largest absolute differences:" % (tag, tot))
    for d, nrm, n in sorted(rows, reverse=True)[:12]:
        print("   %-44s |diff| %.3e  |g| %.3e  rel %.2e" % (n, d, nrm, d / max(nrm, 1e-30)))
    order = list(ga.keys())
    firsts = [n for n in order if np.linalg.norm(ga[n].astype(np.float64) - gb[n].astype(np.float64)) > 1e-6 * max(np.linalg.norm(ga[n]), 1e-30)]
    print("   tensors differing by more than 1e-6 relative: %d of %d; last such in forward order: %s" %
          (len(firsts), len(order), firsts[-1] if firsts else None))


e1 = _engine(variant, nseg, S, F, B, classes, prec, dict(kind="bce", pos_weight=3.0, class_weight=cw))
g1, s1 = run(e1)
g1b, s1b = run(e1)
report("same engine, two steps", g1, g1b)
probe_report("same engine, two steps", ACTS[0], ACTS[1])
fwd_report("same engine, two steps", FWD[0], FWD[1])
pad = torch.empty(64 << 20, dtype=torch.uint8, device="cuda:0").fill_(0xFF)
e2 = _engine(variant, nseg, S, F, B, classes, prec, dict(kind="bce", pos_weight=3.0, class_weight=cw))
g2, s2 = run(e2)
print("logits: same engine %.2e, two engines %.2e" % (np.abs(s1 - s1b).max(), np.abs(s1 - s2).max()))
report("two engines", g1, g2)
