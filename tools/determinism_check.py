"""Runs the same training step twice from the same parameters and reports how far the two gradient vectors differ.

The only run-to-run freedom of a step is the order of floating-point atomics (BN sums in double, weight gradients in
fp32), so loss must agree to ~1e-7 and the gradients to ~1e-5 (norm-wise); anything larger would be a race.

  python tools/determinism_check.py [--precision bf16] [--batch 16] [--size 320] [--trials 3]
"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "instance-segment-basi_b200")):
    sys.path.insert(0, p)

import torch  # noqa: E402

from basi_b200.BAISRunnerTrain import Train  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--precision", default="bf16")
ap.add_argument("--variant", default="1NoClass")
ap.add_argument("--batch", type=int, default=16)
ap.add_argument("--size", type=int, default=320)
ap.add_argument("--trials", type=int, default=3)
a = ap.parse_args()

tr = Train(batch_size=a.batch, last_pool_size=a.size // 8, input_size=[a.size, a.size], log_dir="/tmp/basi_det",
           variant=a.variant, precision=a.precision, use_cuda_graph=False)
eng = tr.engine
img, clicks, lab, cls = tr.data_reader.next_batch()
eng.feed_clicks(img, clicks)
eng.feed(None, lab, cls, 5e-3)
p0 = eng.params_flat.clone()
ref = None
worst = 0.0
for t in range(a.trials):
    eng.params_flat.copy_(p0)
    eng._refresh_weight_copies()
    torch.cuda.synchronize()
    eng.step_device()
    torch.cuda.synchronize()
    g = eng.grads_flat.double().clone()
    loss = eng.losses()[0]
    if ref is None:
        ref, ref_loss = g, loss
        print("trial 0: loss %.10f  |g| %.6e" % (loss, g.norm().item()))
        continue
    rel = ((g - ref).norm() / ref.norm()).item()
    worst = max(worst, rel)
    print("trial %d: loss %.10f (diff %.2e)  |g - g0| / |g0| = %.3e" % (t, loss, abs(loss - ref_loss), rel))
print("worst gradient rel-l2 difference: %.3e" % worst)
