"""Times single tcgen05 conv plans (CUDA events, L2 flushed between launches) for the layer shapes of the benchmark."""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "instance-segment-basi_b200")):
    sys.path.insert(0, p)
import torch  # noqa: E402

from basi_b200 import _lib  # noqa: E402
from basi_b200._lib import ConvDesc  # noqa: E402
from basi_b200.engine import Act  # noqa: E402

SHAPES = [  # name, k, dil, cin, cout, H, B
    ("conv4_1x1_increase", 1, 1, 128, 512, 40, 16),
    ("conv4_1x1_reduce", 1, 1, 512, 128, 40, 16),
    ("conv4_3x3", 3, 2, 128, 128, 40, 16),
    ("conv5_3x3", 3, 4, 256, 256, 40, 16),
    ("conv5_4", 3, 1, 2048, 256, 40, 16),
    ("conv1_3", 3, 1, 32, 64, 160, 16),
    ("conv2_3x3", 3, 1, 32, 32, 80, 16),
    ("x_c64_160", 3, 1, 64, 64, 160, 16),
    ("x_c128_160", 3, 1, 128, 128, 160, 16),
    ("x_1x1_c32_160", 1, 1, 32, 64, 160, 16),
    ("psp_pool1_conv", 1, 1, 1024, 128, 1, 16),
    ("psp_pool6_conv", 1, 1, 1024, 128, 6, 16),
]
if len(sys.argv) > 1:
    SHAPES = [s_ for s_ in SHAPES if s_[0] in sys.argv[1:]]
lib = _lib.load()
st = torch.cuda.current_stream().cuda_stream
flush = torch.zeros(256 << 20, dtype=torch.uint8, device="cuda")
bt = torch.bfloat16


def timeit(fn, n=20, cold=True):
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(n)]
    for _ in range(3):
        fn()
    for a, b in ev:
        if cold:
            flush.zero_()
        a.record()
        fn()
        b.record()
    torch.cuda.synchronize()
    t = sorted(a.elapsed_time(b) for a, b in ev)
    return t[len(t) // 2] * 1e3


for name, k, d, cin, cout, H, B in SHAPES:
    pad = d * (k - 1) // 2
    desc = ConvDesc(k, k, 1, d, pad, pad, 0)
    x = Act(torch.randn((B, H, H, cin), device="cuda").to(bt))
    y = Act(torch.zeros((B, H, H, cout), dtype=bt, device="cuda"))
    dy = Act(torch.randn((B, H, H, cout), device="cuda").to(bt))
    dx = Act(torch.zeros((B, H, H, cin), dtype=bt, device="cuda"))
    w_io = torch.randn(k * k * cin * cout, device="cuda").to(bt)
    w_oi = torch.randn(k * k * cin * cout, device="cuda").to(bt)
    dw = torch.zeros(k * k * cin * cout, device="cuda")
    gam, bet = torch.ones(cout, device="cuda"), torch.zeros(cout, device="cuda")
    sums = torch.zeros(16 * cout, dtype=torch.float64, device="cuda")
    bnp = torch.zeros(4 * cout, device="cuda")
    cnt = torch.zeros(4, dtype=torch.int32, device="cuda")
    flops = 2.0 * B * H * H * cout * k * k * cin
    res = []
    for label, kind, a, b, w, acc, stats in (("fprop", 0, x, y, w_oi, 0, False), ("fprop+stats", 0, x, y, w_oi, 0, True),
                                             ("dgrad", 1, dy, dx, w_io, 0, False), ("dgrad+acc", 1, dy, dx, w_io, 1, False),
                                             ("wgrad", 2, x, dy, None, 1, False)):
        if lib.basi_tc_conv_supported(kind, C.byref(desc), x.ref, y.ref) != 1:
            continue
        h = C.c_void_p()
        _lib.call("basi_tc_conv_create", kind, C.byref(desc), a.ref, b.ref, w.data_ptr() if w is not None else None,
                  dw.data_ptr() if kind == 2 else None, acc, C.byref(h))
        if stats:
            _lib.call("basi_tc_conv_set_bn_stats", h, sums.data_ptr(), gam.data_ptr(), bet.data_ptr(),
                      C.c_double(B * H * H), C.c_float(1e-5), bnp.data_ptr(), cnt.data_ptr())
        us = timeit(lambda: lib.basi_tc_conv_run(h, st))
        uw = timeit(lambda: lib.basi_tc_conv_run(h, st), cold=False)
        res.append("%s %.1f us (%.0f TF/s; warm %.1f us)" % (label, us, flops / us / 1e6, uw))
        lib.basi_tc_conv_destroy(h)
    print("%-20s %s" % (name, " | ".join(res)))
