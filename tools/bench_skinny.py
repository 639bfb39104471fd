"""Micro-benchmark of the attention-class head GEMVs (class_attention_conv: M = batch 16, K = 25 * 1024, N = 512; the
float32 weight matrix is 52 MB and is streamed once per call).  Four weight / gradient buffers are used round-robin
(208 MB > the 126 MB L2), CUDA events around 20 calls each.

  python tools/bench_skinny.py [--M 16] [--K 25600] [--N 512]
"""
import argparse
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "instance-segment-basi_b200")):
    sys.path.insert(0, p)

import torch  # noqa: E402

from basi_b200 import _lib  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--M", type=int, default=16)
ap.add_argument("--K", type=int, default=25600)
ap.add_argument("--N", type=int, default=512)
ap.add_argument("--reps", type=int, default=20)
a = ap.parse_args()
M, K, N = a.M, a.K, a.N
dev = "cuda:0"
st = torch.cuda.current_stream().cuda_stream
ws = [torch.randn(K, N, device=dev) / K ** 0.5 for _ in range(4)]
dws = [torch.zeros(K, N, device=dev) for _ in range(4)]
x = torch.randn(M, K, device=dev).to(torch.bfloat16)
dx = torch.zeros(M, K, device=dev, dtype=torch.bfloat16)
bias, dbias = torch.zeros(N, device=dev), torch.zeros(N, device=dev)
y, dy = torch.zeros(M, N, device=dev), torch.randn(M, N, device=dev)
nws = int(_lib.load().basi_skinny_fwd_workspace_floats(M, K, N))
wsp = torch.zeros(nws, device=dev)
wbytes = K * N * 4


def timed(fn, nbytes, label):
    for i in range(4):
        fn(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(a.reps):
        fn(i)
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / a.reps * 1e3
    print("%-22s %8.1f us   %7.0f GB/s (algorithmic bytes %.1f MB)" % (label, us, nbytes / us / 1e3, nbytes / 1e6))


timed(lambda i: _lib.call("basi_skinny_fwd_ws", x.data_ptr(), 1, C.c_int64(K), ws[i % 4].data_ptr(), bias.data_ptr(),
                          y.data_ptr(), M, K, N, 1, wsp.data_ptr(), st), wbytes, "skinny_fwd_ws")
timed(lambda i: _lib.call("basi_skinny_dgrad", dy.data_ptr(), ws[i % 4].data_ptr(), dx.data_ptr(), 1, C.c_int64(K), M,
                          K, N, 0, st), wbytes, "skinny_dgrad")
timed(lambda i: _lib.call("basi_skinny_wgrad", x.data_ptr(), 1, C.c_int64(K), dy.data_ptr(), dws[i % 4].data_ptr(),
                          dbias.data_ptr(), M, K, N, st), 2 * wbytes, "skinny_wgrad (RMW)")
