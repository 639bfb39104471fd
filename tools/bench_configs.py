"""Throughput of the other BASELINE.json configurations on ONE GPU (their per-GPU share), through the public Train API
with the CUDA-graph step: not bench.py lines (those are cfg2 only), just evidence that the same kernels run the other
variants and resolutions.

  python tools/bench_configs.py [--steps 15]
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "instance-segment-basi_b200")):
    sys.path.insert(0, p)

import torch  # noqa: E402

from basi_b200.BAISRunnerTrain import Train  # noqa: E402

CONFIGS = [  # name, variant, size, per-GPU batch, classes
    ("cfg2 joint training (2AddClass, 0.2*loss_classes), 320x320, batch 16/GPU", "2AddClass", 320, 16, 21),
    ("cfg3 4BorderClass head, 512x512, batch 4/GPU (global 32 on 8 GPUs)", "4BorderClass", 512, 4, 21),
    ("cfg4 5COCO joint training, 640x640, batch 16/GPU (global 128 on 8 GPUs)", "5COCO", 640, 16, 81),
]
ap = argparse.ArgumentParser()
ap.add_argument("--steps", type=int, default=15)
a = ap.parse_args()
for name, variant, size, batch, classes in CONFIGS:
    tr = Train(batch_size=batch, last_pool_size=size // 8, input_size=[size, size], log_dir="/tmp/basi_cfg",
               variant=variant, num_classes=classes, precision="bf16", use_cuda_graph=True)
    for i in range(4):
        tr.run_step(i, fetch=False)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    for i in range(a.steps):
        tr.run_step(4 + i, fetch=False)          # host batch -> pinned H2D -> graph replay, every step
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / a.steps
    loss = tr.engine.losses()
    print(json.dumps({"config": name, "images_per_s": batch / (ms * 1e-3), "ms_per_step": ms, "loss": loss[0],
                      "tc_layers": tr.engine.tc_layers, "launches_per_step": tr.engine.launches_per_step()}))
    del tr
    torch.cuda.empty_cache()
