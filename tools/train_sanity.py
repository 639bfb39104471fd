"""Short synthetic training run (Train.run_step) in one precision: prints the loss every N steps and checks that it
stays finite -- the range check of the fp16 mode's static loss scale.

  python tools/train_sanity.py --precision f16 --steps 300 [--variant 2AddClass] [--batch 16]
"""
import argparse
import math
import os
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
ap.add_argument("--steps", type=int, default=300)
ap.add_argument("--every", type=int, default=25)
a = ap.parse_args()
tr = Train(batch_size=a.batch, last_pool_size=a.size // 8, input_size=[a.size, a.size], log_dir="/tmp/basi_sanity",
           variant=a.variant, precision=a.precision, num_steps=100000)
losses = []
for step in range(a.steps):
    r = tr.run_step(step, fetch=(step % a.every == 0 or step == a.steps - 1))
    if r is not None:
        losses.append(r["loss"])
        print("step %4d  loss %.5f  seg %.5f  cls %.5f" % (step, r["loss"], r["loss_segment"], r["loss_classes"]), flush=True)
        assert math.isfinite(r["loss"]), "loss is not finite"
g = tr.engine.grads_flat
print("max |scaled grad| %.3e (loss scale %g), non-finite gradient entries: %d" % (
    float(g.abs().max()), tr.engine.loss_scale, int((~torch.isfinite(g)).sum())))
print("first %.4f last %.4f" % (losses[0], losses[-1]))
