#!/usr/bin/env python
"""Benchmark of the BAIS PSPNet hot path: train images/sec at 320^2 (BASELINE.json metric).

  python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path (N>1: under torchrun)
  python bench.py --impl reference --gpus N --steps K ...  # the CPU restatement of the reference (oracle)

Workload = BASELINE.json configs[1]: segment-only training (1NoClass head, pos_weight=3), synthetic
VOC-shaped 320x320 inputs, batch 16 per GPU, random-init weights; N>1 is weak scaling (16 per GPU),
data-parallel with a bucketed NCCL gradient all-reduce.  One step = click-map pack + forward + fused loss
+ backward + SGD.  Prints ONE JSON line (rank 0).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "instance-segment-basi_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)

METRIC = "train images/sec at 320^2"
UNIT = "images/s"
S, P, BATCH, FILTERS = 320, 40, 16, 32
VARIANT, NSEG, CLASSES = "1NoClass", 1, 21
FLOP_PER_IMAGE = 153.3e9       # fwd+bwd conv FLOPs per image at 320^2


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        d = json.load(open(path))
        return dict(hbm=d["hbm_gbs"], tensor=d["bf16_tflops"], tensor_sustained=d["bf16_tflops_sustained"],
                    source="measured")
    return dict(hbm=6650.0, tensor=1590.0, tensor_sustained=1400.0, source="fallback")


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    FIELDS = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        threading.Thread.__init__(self, daemon=True)
        self.index, self.samples, self.stop_flag, self.proc = index, [], False, None

    def run(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.FIELDS,
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            for line in self.proc.stdout:
                if self.stop_flag:
                    break
                self.samples.append([x.strip() for x in line.split(",")])
        except Exception:
            pass

    def stop(self):
        self.stop_flag = True
        if self.proc is not None:
            self.proc.terminate()

    def summary(self):
        sm, mx, reasons = [], 0.0, set()
        names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
        for s in self.samples:
            try:
                sm.append(float(s[0]))
                mx = max(mx, float(s[1]))
            except Exception:
                continue
            for n, v in zip(names, s[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx or None, "reasons": sorted(reasons),
                "samples": len(sm)}


def cpu_oracle_rate(batch, steps, warmup=1):
    """Oracle (CPU restatement of the reference, TF unavailable) timed on the host cores: images/s."""
    import numpy as np
    import torch
    from basi_b200.BAISData import SyntheticData
    from oracle import basi_oracle as O
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    sd = SyntheticData(batch, (S, S), 8, CLASSES, NSEG, seed=0)
    params = O.init_params(O.param_specs(VARIANT, CLASSES, NSEG, FILTERS), 0)
    times = []
    for i in range(warmup + steps):
        img, clicks, lab, cls = sd.next_batch()
        t0 = time.perf_counter()
        data = np.stack([O.pack_input(img[b], clicks[b]) for b in range(batch)])
        r = O.train_step(params, data, lab, cls, VARIANT, NSEG, P, 3.0, 0.0, 5e-3, torch.float32)
        params = r["new_params"]
        dt = time.perf_counter() - t0
        if i >= warmup:
            times.append(dt)
    return batch * len(times) / sum(times), cores, sum(times) / len(times)


def click_latency(torch, precision, n=30):
    """cfg 1: single-click inference (RunnerGUI semantics, 4BorderClass head, S=320, B=1): host image + click in,
    thresholded full-resolution mask + class out.  p50 / p90 wall-clock latency in ms."""
    import numpy as np
    from basi_b200.BAISRunnerOne import RunnerGUI
    gui = RunnerGUI(None, last_pool_size=P, variant="4BorderClass", num_classes=21, num_segment=4,
                    filter_number=FILTERS, precision=precision)
    rng = np.random.RandomState(0)
    img = rng.randint(0, 256, size=(S, S, 3), dtype=np.uint8)
    for _ in range(3):
        gui.click(img, [160, 160])
    ts = []
    for i in range(n):
        where = [int(rng.randint(0, S)), int(rng.randint(0, S))]
        t0 = time.perf_counter()
        gui.click(img, where)
        ts.append((time.perf_counter() - t0) * 1e3)
    ts.sort()
    return {"p50_ms": ts[len(ts) // 2], "p90_ms": ts[int(len(ts) * 0.9)], "n": n,
            "h2d_bytes": S * S * 3 + 8, "d2h_bytes": S * S * 4 + 4,
            "workload": "cfg1: RunnerGUI single click, 320x320, batch 1, 4BorderClass head, CUDA-graph forward"}


def cpu_click_latency(n=3):
    import numpy as np
    import torch
    from oracle import basi_oracle as O
    torch.set_num_threads(os.cpu_count() or 1)
    params = O.to_torch(O.init_params(O.param_specs("4BorderClass", 21, 4, FILTERS), 0), torch.float32)
    rng = np.random.RandomState(0)
    img = rng.randint(0, 256, size=(S, S, 3), dtype=np.uint8)
    ts = []
    with torch.no_grad():
        for i in range(n + 1):
            t0 = time.perf_counter()
            data = O.pack_input(img, [160, 160])[None]
            out = O.pspnet_forward(params, torch.from_numpy(data), "4BorderClass", 4, P)
            O.predict_click(out["conv6_n_4"].numpy(), (S, S))
            if i:
                ts.append((time.perf_counter() - t0) * 1e3)
    ts.sort()
    return ts[len(ts) // 2]


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    batch = 2
    rate, cores, sec = cpu_oracle_rate(batch, max(1, args.steps), 1)
    sample = "oracle fwd+bwd+SGD, S=320, batch %d per step, float32 torch-CPU (%d threads)" % (batch, cores)
    line = {"impl": "reference", "metric": METRIC, "value": rate, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "cfg2: segment-only training (1NoClass, pos_weight=3), 320x320, CPU sample batch 2",
                       "note": "reference TF1 cannot run here; this is the CPU oracle restating it"},
            "cpu_baseline": {"value": rate, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": rate, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def kernel_profile(eng, torch, detail_path=None):
    """One instrumented eager step: CUDA events around every C-ABI call, aggregated per kernel entry point."""
    st = torch.cuda.current_stream()
    eng._zero_step_state(st.cuda_stream)
    agg = {}
    detail = []
    for lst in (eng.pre, eng.fwd, eng.lossl, eng.post, eng.bwd):
        evs = [torch.cuda.Event(enable_timing=True) for _ in range(len(lst) + 1)]
        evs[0].record(st)
        for i, (name, fn, a, meta) in enumerate(lst):
            rc = fn(*a, st.cuda_stream)
            assert rc == 0, name
            evs[i + 1].record(st)
        torch.cuda.synchronize()
        for i, (name, fn, a, meta) in enumerate(lst):
            ms = evs[i].elapsed_time(evs[i + 1])
            # fprop (:0) and dgrad (:1) plans run the same kernel, conv_tc_kernel; wgrad (:2) is wgrad_tc_kernel
            name = {"basi_tc_conv_run:0": "conv_tc_kernel(fprop+dgrad)", "basi_tc_conv_run:1": "conv_tc_kernel(fprop+dgrad)",
                    "basi_tc_conv_run:2": "wgrad_tc_kernel"}.get(name, name)
            d = agg.setdefault(name, dict(ms=0.0, n=0, flops=0.0, bytes=0.0))
            d["ms"] += ms
            d["n"] += 1
            d["flops"] += meta.get("flops", 0.0)
            d["bytes"] += meta.get("bytes", 0.0)
            detail.append((name, meta.get("layer", ""), ms, meta.get("flops", 0.0), meta.get("bytes", 0.0)))
    if detail_path:
        with open(detail_path, "w") as f:
            for name, layer, ms, fl, by in detail:
                f.write("%s\t%s\t%.2f us\t%s\n" % (name, layer, ms * 1e3, ("%.0f TF/s" % (fl / ms / 1e9)) if fl else
                                                     (("%.0f GB/s" % (by / ms / 1e6)) if by else "")))
    return agg


def run_ours(args):
    import numpy as np
    import torch
    from basi_b200 import _lib
    from basi_b200.BAISRunnerTrain import Train
    from basi_b200.dp import DataParallel

    world = int(os.environ.get("WORLD_SIZE", "1"))
    os.environ.pop("NCCL_DEBUG", None)         # keep NCCL's version banner off stdout (one JSON line only)
    dp = DataParallel() if world > 1 else None
    rank = dp.rank if dp else 0
    dev_index = dp.local_rank if dp else 0
    torch.cuda.set_device(dev_index)
    device = "cuda:%d" % dev_index
    tr = Train(batch_size=BATCH, last_pool_size=P, input_size=[S, S], log_dir="/tmp/basi_bench_%d" % rank,
               variant=VARIANT, num_classes=CLASSES, precision=args.precision, filter_number=FILTERS, seed=0,
               device=device, dp=dp, use_cuda_graph=not args.no_graph, use_tc=not args.no_tc)
    eng = tr.engine
    if dp:
        dp.broadcast(eng.params_flat)
    sd = tr.data_reader
    # ---- pinned host batches (a small ring, refilled round-robin)
    ring = []
    for _ in range(4):
        img, clicks, lab, cls = sd.next_batch()
        ring.append((torch.from_numpy(img).pin_memory(), torch.from_numpy(clicks).pin_memory(),
                     torch.from_numpy(np.ascontiguousarray(lab, dtype=np.float32)).pin_memory(), cls))
    h2d = int(ring[0][0].numel() + ring[0][1].numel() * 4 + ring[0][2].numel() * 4 + 4)
    d2h = 8 * 4

    def barrier():
        torch.cuda.synchronize()
        if dp:
            dp.barrier()
        torch.cuda.synchronize()

    def device_step():
        if tr.use_cuda_graph:
            eng.replay()
        elif dp:
            eng.step_device(sync_grads=dp)
        else:
            eng.step_device()

    # warm-up (also captures the CUDA graph)
    eng.feed_clicks(ring[0][0], ring[0][1])
    eng.feed(None, ring[0][2], None, 5e-3)
    if tr.use_cuda_graph:
        try:
            eng.capture(train=True, sync_grads=dp)
        except Exception as e:        # e.g. an NCCL build that cannot be captured: fall back to eager launches
            if not dp:
                raise
            sys.stderr.write("CUDA-graph capture with NCCL failed (%s); running eagerly\n" % (e,))
            tr.use_cuda_graph = False
            torch.cuda.synchronize()
    for _ in range(max(3, args.warmup)):
        device_step()
    barrier()

    sampler = ClockSampler(dev_index) if rank == 0 else None
    if sampler:
        sampler.start()
        time.sleep(0.3)
    # ---- (1) device-resident throughput: inputs already in HBM, K steps between events
    launches0 = _lib.LAUNCHES
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(args.steps):
        device_step()
    e1.record()
    barrier()
    ms_dev = e0.elapsed_time(e1)
    launches = _lib.LAUNCHES - launches0
    # ---- (2) end to end through Train.run_step-equivalent: pinned H2D of every step's inputs, D2H of the loss
    barrier()
    t0 = time.perf_counter()
    e0.record()
    last_loss = None
    for i in range(args.steps):
        img, clicks, lab, cls = ring[i % len(ring)]
        eng.feed_clicks(img, clicks)
        eng.feed(None, lab, None, 5e-3)
        device_step()
        last_loss = eng.losses()          # D2H read of the step's loss (synchronises)
    e1.record()
    barrier()
    ms_e2e = e0.elapsed_time(e1)
    if sampler:
        sampler.stop()
    t = torch.tensor([ms_dev, ms_e2e], dtype=torch.float64, device=device)
    if dp:
        import torch.distributed as dist
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_dev, ms_e2e = float(t[0]), float(t[1])
    if rank != 0:
        return
    pk = peaks()
    images = world * BATCH * args.steps
    value = images / (ms_dev * 1e-3)
    e2e = images / (ms_e2e * 1e-3)
    # ---- roofline of the dominant kernel (per-launch CUDA-event times of one eager step)
    agg = kernel_profile(eng, torch, args.detail)
    total_ms = sum(d["ms"] for d in agg.values())
    top = max(agg.items(), key=lambda kv: kv[1]["ms"])
    name, d = top
    if d["flops"] > 0:
        ach = d["flops"] / (d["ms"] * 1e-3) / 1e12
        roof = {"bound": "tensor", "kernel": name, "achieved": ach, "peak": pk["tensor_sustained"], "unit": "TFLOP/s",
                "frac": ach / pk["tensor_sustained"], "traffic": None, "peak_source": pk["source"] + " sustained bf16",
                "launches_per_step": d["n"], "share_of_step": d["ms"] / total_ms}
    else:
        ach = d["bytes"] / (d["ms"] * 1e-3) / 1e9
        roof = {"bound": "hbm", "kernel": name, "achieved": ach, "peak": pk["hbm"], "unit": "GB/s",
                "frac": ach / pk["hbm"], "traffic": None, "peak_source": pk["source"],
                "launches_per_step": d["n"], "share_of_step": d["ms"] / total_ms}
    # DRAM bytes per launch of that kernel group from the committed ncu capture of the same workload
    # (profiles/traffic_r01.json, made by tools/make_traffic.py from the launch list of tools/profile_step.py)
    try:
        with open(os.path.join(ROOT, "profiles", "traffic_r01.json")) as f:
            grp = json.load(f).get("groups", {}).get(name)
        if grp:
            roof["traffic"] = grp["dram_bytes_per_launch"]
            roof["traffic_source"] = "ncu dram__bytes_read.sum + dram__bytes_write.sum per launch, profiles/traffic_r01.json"
    except (OSError, ValueError):
        pass
    breakdown = {k: {"ms": round(v["ms"], 3), "n": v["n"],
                     "tflops": round(v["flops"] / (v["ms"] * 1e-3) / 1e12, 2) if v["flops"] else None,
                     "gbs": round(v["bytes"] / (v["ms"] * 1e-3) / 1e9, 1) if v["bytes"] else None}
                 for k, v in sorted(agg.items(), key=lambda kv: -kv[1]["ms"])[:12]}
    # ---- CPU baseline (bounded sample on the host cores)
    cpu = None
    if not args.no_cpu:
        rate, cores, sec = cpu_oracle_rate(2, 2, 1)
        cpu = {"value": rate, "unit": UNIT, "cores": cores, "kind": "port",
               "sample": "oracle fwd+bwd+SGD at S=320, batch 2 x 2 steps, float32 torch-CPU (%.1f s/step)" % sec}
    use_graph, tc_layers = bool(tr.use_cuda_graph), eng.tc_layers
    del tr, eng
    torch.cuda.empty_cache()
    click = None
    if world == 1 and not args.no_click:
        click = click_latency(torch, args.precision)
        if not args.no_cpu:
            click["cpu_oracle_p50_ms"] = cpu_click_latency()
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(3, args.warmup), "ms_per_step": ms_dev / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": args.precision, "data": "synthetic",
            "config": {"workload": "cfg2: BAISNet segment-only training (1NoClass head, pos_weight=3), synthetic "
                                   "VOC-shape 320x320, batch 16 per GPU, F=32, SGD",
                       "global_batch": world * BATCH, "parallelism": "dp%d" % world,
                       "l2": "per-step working set (~3 GB of activations) exceeds the 126 MB L2",
                       "cuda_graph": use_graph, "tc_layers": tc_layers,
                       "model_tflops": value * FLOP_PER_IMAGE / 1e12},
            "e2e": {"value": e2e, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": ms_e2e / args.steps, "loss": last_loss[0] if last_loss else None},
            "gpu_launches": int(launches),
            "roofline": roof, "kernel_breakdown_ms_per_step": breakdown, "cpu_baseline": cpu,
            "click_to_mask": click,
            "clocks": sampler.summary() if sampler else None}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--precision", default="bf16", choices=["bf16", "f32"])
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--no-tc", action="store_true")
    ap.add_argument("--no-cpu", action="store_true", help="skip the CPU-oracle baseline leg")
    ap.add_argument("--no-click", action="store_true", help="skip the click-to-mask latency leg (cfg 1)")
    ap.add_argument("--detail", default=None, help="write per-call CUDA-event timings of one eager step to this file")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)
    try:
        import torch.distributed as dist
        if dist.is_initialized():
            dist.destroy_process_group()
    except Exception:
        pass


if __name__ == "__main__":
    main()
