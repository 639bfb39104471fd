"""Where does the in-graph step time go?  Captures the benchmark step with groups of calls REMOVED from the plan and
times the graph replay (results of an ablated step are garbage; only the time is read).  The marginal time of a group
inside the real CUDA graph -- with programmatic dependent launch and the side stream overlapping -- is what an
optimisation of that group can win, unlike the per-launch CUDA-event sums of `bench.py --detail`.

  python tools/ablate_step.py [--variant 2AddClass] [--precision f16] [--batch 16] [--steps 30]
"""
import argparse
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "instance-segment-basi_b200")):
    sys.path.insert(0, p)

import torch  # noqa: E402

from basi_b200.BAISRunnerTrain import Train  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--precision", default="f16")
ap.add_argument("--variant", default="2AddClass")
ap.add_argument("--batch", type=int, default=16)
ap.add_argument("--size", type=int, default=320)
ap.add_argument("--steps", type=int, default=30)
ap.add_argument("--only", default="")
a = ap.parse_args()

# name -> (which lists, predicate on the call name) of the calls to DROP
ABLATIONS = [
    ("full step", None),
    ("no wgrad (tcgen05 + SIMT)", lambda n, m: "wgrad" in n or n.startswith("basi_tc_conv_run:2")),
    ("no BN backward", lambda n, m: n.startswith("basi_bn_bwd")),
    ("no BN backward, no wgrad", lambda n, m: n.startswith("basi_bn_bwd") or "wgrad" in n or n.startswith("basi_tc_conv_run:2")),
    ("no dgrad", lambda n, m: n.startswith("basi_tc_conv_run:1") or "dgrad" in n),
    ("no BN apply (forward)", lambda n, m: n.startswith("basi_bn_apply") or n.startswith("basi_bnact")),
    ("no fprop convs", lambda n, m: n.startswith("basi_tc_conv_run:0") or n.startswith("basi_conv_fprop")),
    ("forward only", "fwd"),
    ("backward lists empty + no sgd", "nobwd"),
]


def build():
    tr = Train(batch_size=a.batch, last_pool_size=a.size // 8, input_size=[a.size, a.size], log_dir="/tmp/basi_abl",
               variant=a.variant, precision=a.precision, use_cuda_graph=True)
    eng = tr.engine
    img, clicks, lab, cls = tr.data_reader.next_batch()
    eng.feed_clicks(img, clicks)
    eng.feed(None, lab, cls, 5e-3, **({"label_att": tr._attention_labels(lab)} if a.variant == "8AttentionU" else {}))
    return tr, eng


def timed(eng, train=True):
    eng._graph = None
    eng.capture(train=train)
    for _ in range(3):
        eng.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.steps):
        eng.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / a.steps


tr, eng = build()
names = sorted(set(re.sub(r":\d+$", lambda m: m.group(0), c[0]) for c in eng.fwd + eng.bwd + eng.lossl))
print("call kinds:", ", ".join(names))
saved = dict(fwd=list(eng.fwd), bwd=list(eng.bwd), lossl=list(eng.lossl))
base = None
for label, pred in ABLATIONS:
    if a.only and a.only not in label:
        continue
    eng.fwd, eng.bwd, eng.lossl = list(saved["fwd"]), list(saved["bwd"]), list(saved["lossl"])
    train = True
    if pred == "fwd":
        train = False
    elif pred == "nobwd":
        eng.bwd = []
    elif pred is not None:
        eng.fwd = [c for c in eng.fwd if not pred(c[0], c[3])]
        eng.bwd = [c for c in eng.bwd if not pred(c[0], c[3])]
    n = len(eng.fwd) + (len(eng.bwd) + len(eng.lossl) if train else 0)
    ms = timed(eng, train)
    base = ms if base is None else base
    print("%-34s %4d calls  %7.3f ms/step  (%+.3f ms vs full)" % (label, n, ms, ms - base))
