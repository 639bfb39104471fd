"""Experiment: scheduling of the tensor-core weight gradients (BASI_WGRAD_STREAMS / BASI_DEFER_WGRAD).  Times the
CUDA-graph step of the benchmark workload per mode and checks that the gradients agree with the default schedule.

  BASI_EXPERIMENTS=1 python tools/wgrad_sched.py [--variant 2AddClass] [--precision f16]
"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "instance-segment-basi_b200")):
    sys.path.insert(0, p)
os.environ["BASI_EXPERIMENTS"] = "1"

import numpy as np  # noqa: E402
import torch  # noqa: E402

from basi_b200.BAISRunnerTrain import Train  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--precision", default="f16")
ap.add_argument("--variant", default="2AddClass")
ap.add_argument("--batch", type=int, default=16)
ap.add_argument("--steps", type=int, default=30)
ap.add_argument("--modes", default="1:0,2:0,3:0,1:1,2:1,3:1,4:1,6:1",
                help="comma list of streams:defer[:wgrad_splits[:wgrad_waves[:wgrad_target_ctas[:maxtiles[:stages[:sm_main[:sm_wgrad]]]]]]]")
a = ap.parse_args()

ref = None
for mode in a.modes.split(","):
    parts = mode.split(":") + ["", "", "", "", "", "", ""]
    k, d, sp, wv, tg, mx, sg = parts[0], parts[1], parts[2], parts[3], parts[4], parts[5], parts[6]
    sm_main, sm_wg = parts[7], parts[8]
    for name, val in (("BASI_TC_WGRAD_SPLITS", sp), ("BASI_TC_WGRAD_WAVES", wv), ("BASI_TC_WGRAD_TARGET", tg),
                      ("BASI_TC_WGRAD_MAXTILES", mx), ("BASI_TC_WGRAD_STAGES", sg), ("BASI_SM_MAIN", sm_main),
                      ("BASI_SM_WGRAD", sm_wg)):
        if val:
            os.environ[name] = val
        else:
            os.environ.pop(name, None)
    os.environ["BASI_WGRAD_STREAMS"] = k
    if d == "1":
        os.environ["BASI_DEFER_WGRAD"] = "1"
    else:
        os.environ.pop("BASI_DEFER_WGRAD", None)
    tr = Train(batch_size=a.batch, last_pool_size=40, input_size=[320, 320], log_dir="/tmp/basi_ws", variant=a.variant,
               precision=a.precision, use_cuda_graph=True, seed=0)
    eng = tr.engine
    from basi_b200.BAISData import SyntheticData
    img, clicks, lab, cls = SyntheticData(a.batch, (320, 320), 8, 21, tr.num_segment, seed=5).next_batch()
    eng.feed_clicks(img, clicks)
    eng.feed(None, lab, cls, 0.0)
    eng.step_device()
    torch.cuda.synchronize()
    g = eng.grads_flat.float().cpu().numpy().copy()
    if ref is None:
        ref = g
    err = float(np.linalg.norm(g - ref) / np.linalg.norm(ref))
    eng.capture(train=True)
    for _ in range(3):
        eng.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.steps):
        eng.replay()
    e1.record()
    torch.cuda.synchronize()
    print("wgrad streams %s defer %s splits %s waves %s target %s maxtiles %s stages %s sm_main %s sm_wgrad %s: %.3f ms/step, "
          "gradient rel-l2 vs first mode %.2e" % (k, d, sp or "-", wv or "-", tg or "-", mx or "-", sg or "-",
                                                  sm_main or "-", sm_wg or "-", e0.elapsed_time(e1) / a.steps, err),
          flush=True)
    del tr, eng
    torch.cuda.empty_cache()
