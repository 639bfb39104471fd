"""One eager training step of the benchmark workload between cudaProfilerStart/Stop (for ncu --profile-from-start off).

  python tools/profile_step.py [--precision bf16] [--variant 1NoClass] [--batch 16] [--no-tc]
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
ap.add_argument("--no-tc", action="store_true")
ap.add_argument("--steps", type=int, default=1)
a = ap.parse_args()

tr = Train(batch_size=a.batch, last_pool_size=a.size // 8, input_size=[a.size, a.size], log_dir="/tmp/basi_prof",
           variant=a.variant, precision=a.precision, use_cuda_graph=False, use_tc=not a.no_tc)
eng = tr.engine
img, clicks, lab, cls = tr.data_reader.next_batch()
eng.feed_clicks(img, clicks)
eng.feed(None, lab, cls, 5e-3)
for _ in range(2):
    eng.step_device()
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
for _ in range(a.steps):
    eng.step_device()
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
print("profiled %d step(s); loss %s; tc plans %d" % (a.steps, eng.losses(), eng.tc_layers))
