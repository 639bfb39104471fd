"""CPU, world_size 2 over gloo: the data-parallel exchange (bucket plan + averaged all-reduce)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

from basi_b200.BAISPSPNet import PSPNet, Placeholder
from basi_b200.dp import DataParallel
from basi_b200.engine import Engine


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out):
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1",
                      MASTER_PORT=str(port))
    dp = DataParallel(backend="gloo")
    flat = torch.arange(10, dtype=torch.float32) * (rank + 1)
    dp.all_reduce_mean(flat)
    w = torch.full((4,), float(rank))
    dp.broadcast(w, 0)
    dp.barrier()
    out.put((rank, flat.numpy().tolist(), w.numpy().tolist()))
    torch.distributed.destroy_process_group()


def test_all_reduce_mean_and_broadcast_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    ps = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in ps:
        p.start()
    res = sorted(q.get(timeout=120) for _ in ps)
    for p in ps:
        p.join(60)
        assert p.exitcode == 0
    want = (np.arange(10) * 1.5).tolist()
    for rank, flat, w in res:
        assert flat == want          # mean of (x, 2x)
        assert w == [0.0, 0.0, 0.0, 0.0]


def test_bucket_plan_covers_flat_buffer_in_reverse_order():
    net = PSPNet({'data': Placeholder((None, 320, 320, 4))}, num_classes=21, num_segment=1, is_training=True,
                 last_pool_size=40, filter_number=32)
    e = Engine(net, 2, "bf16", True, dict(kind="bce", pos_weight=3.0, class_weight=0.2), dry_run=True)
    plan = DataParallel.plan_buckets(e.param_index, e.bwd, e.n_flat, (25 << 20) // 4)
    # contiguous cover of [0, n_flat), cut from the end
    assert plan[0][2] == e.n_flat and plan[-1][1] == 0
    for (r0, s0, e0), (r1, s1, e1) in zip(plan, plan[1:]):
        assert e1 == s0 and r1 >= r0            # launch order never goes backwards
    assert all(0 <= r < len(e.bwd) for r, _, _ in plan)
    # the class head's 13.1 M-parameter conv finishes first and sits in the first bucket
    off = e.param_index["class_attention_conv/weights"][0]
    assert plan[0][1] <= off < plan[0][2]
    # every bucket is complete once its ready index has run: no later call writes into it
    for ready, start, end in plan:
        for i, call in enumerate(e.bwd):
            for name in call[3].get("writes", ()):
                o = e.param_index[name][0]
                if start <= o < end:
                    assert i <= ready, (name, i, ready)
