"""Multi-GPU check of the gradient exchange (run under torchrun, one process per GPU):

  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/dp_check.py

Every rank runs the same small network on its own batch; the averaged gradient of the peer-memory all-reduce
(csrc/p2p.cu) must equal an NCCL all-reduce (AVG) of the local gradients bit for bit up to the float32 summation order,
on every rank, over several steps (the flag barriers are reused)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "instance-segment-basi_b200")):
    sys.path.insert(0, p)

import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from basi_b200.BAISData import SyntheticData  # noqa: E402
from basi_b200.BAISPSPNet import PSPNet, Placeholder  # noqa: E402
from basi_b200.dp import DataParallel  # noqa: E402
from basi_b200.engine import Engine  # noqa: E402

dp = DataParallel()
dev = "cuda:%d" % dp.local_rank
S, F, B = 64, 16, 2
net = PSPNet({'data': Placeholder((None, S, S, 4))}, num_classes=21, num_segment=1, is_training=True,
             last_pool_size=S // 8, filter_number=F, variant="2AddClass")
eng = Engine(net, B, "f16", True, dict(kind="bce", pos_weight=3.0, class_weight=0.2), dev)
eng.init_params(0)
eng.broadcast_params(dp)
eng.enable_click_input(30)
sd = SyntheticData(B, (S, S), 8, 21, 1, seed=100 + dp.rank)
worst = 0.0
st = torch.cuda.current_stream().cuda_stream
for step in range(4):
    img, clicks, lab, cls = sd.next_batch()
    eng.feed_clicks(img, clicks)
    eng.feed(None, lab, cls, 0.0)                      # lr 0: the weights stay equal on all ranks
    eng.step_device(sync_grads=dp)                     # a whole step through the exchange (sets the p2p path up)
    torch.cuda.synchronize()
    assert torch.isfinite(eng.grads_flat).all(), "gradient is not finite"
    # the exchange itself, on identical inputs: peer-memory all-reduce vs NCCL AVG
    local = torch.randn(eng.n_flat, device=dev) * (1 + dp.rank) + eng.grads_flat
    ref = local.clone()
    dist.all_reduce(ref, op=dist.ReduceOp.AVG)
    got = local.clone()
    if dp.p2p is not None:
        dp.p2p.all_reduce_mean(got, st)
    else:
        dp.all_reduce_mean(got)
    torch.cuda.synchronize()
    worst = max(worst, float((got - ref).abs().max() / ref.abs().max()))
t = torch.tensor([worst], device=dev)
dist.all_reduce(t, op=dist.ReduceOp.MAX)
if dp.rank == 0:
    print("exchange: %s; max relative difference to NCCL AVG over 4 steps and %d ranks: %.3e" % (
        "p2p" if dp.p2p is not None else "nccl", dp.world, float(t)))
    assert float(t) < 1e-6
dist.destroy_process_group()
