"""Data parallelism for the training step: one process per GPU, NCCL gradient all-reduce over NVLink.

The reference is single-GPU (os.environ["CUDA_VISIBLE_DEVICES"], back/2AddClass/BAISRunnerTrain.py);
the path shards by samples: every rank holds a full replica and its own batch slice,
batch-norm statistics stay per replica (== the reference at batch B/N), and the only exchange is one
averaged all-reduce of the flat gradient buffer per step.  The buffer is cut into buckets in reverse
parameter order; each bucket's all-reduce is issued on a side stream as soon as the last backward call
that writes into it has been enqueued, so the transfer overlaps the remaining backward kernels.
"""
from __future__ import annotations

import ctypes as C
import os
import sys

import torch
import torch.distributed as dist


class P2PAllReduce(object):
    """Averaged all-reduce of the flat gradient over NVLink peer memory (csrc/p2p.cu: flag barrier, every rank
    averages its slice of all ranks' buffers and writes it back to all of them, flag barrier -- one full-machine
    kernel per rank instead of NCCL kernels that the persistent backward kernels starve).  The symmetric buffer and
    the peer mappings come from torch's symmetric memory (plumbing); the exchange itself is ours."""

    def __init__(self, rank, world, n_floats, device):
        import torch.distributed._symmetric_memory as symm_mem
        from . import _lib
        assert n_floats % 4 == 0
        self.rank, self.world, self.n = rank, world, int(n_floats)
        group = dist.group.WORLD
        try:
            symm_mem.enable_symm_mem_for_group(group.group_name)
        except Exception:
            pass
        self.buf = symm_mem.empty(self.n + 64, dtype=torch.float32, device=device)
        self.buf.zero_()
        self.hdl = symm_mem.rendezvous(self.buf, group)
        ptrs = [int(p) for p in self.hdl.buffer_ptrs]
        if len(ptrs) != world or self.hdl.rank != rank:
            raise RuntimeError("symmetric memory rendezvous returned %d buffers for world %d" % (len(ptrs), world))
        self._bufs = (C.c_void_p * world)(*ptrs)
        self._flags = (C.c_void_p * world)(*[p + 4 * self.n for p in ptrs])     # uint32[world] behind the gradients
        self.seq = torch.zeros(1, dtype=torch.int32, device=device)
        self._lib = _lib
        torch.cuda.synchronize(device)
        dist.barrier()                       # every rank's flags are zero before anyone signals

    def all_reduce_mean(self, flat, stream):
        self.buf[:self.n].copy_(flat, non_blocking=True)
        self._lib.call("basi_p2p_allreduce_mean", self._bufs, self._flags, self.rank, self.world, C.c_int64(self.n),
                       self.seq.data_ptr(), stream)
        flat.copy_(self.buf[:self.n], non_blocking=True)


class DataParallel(object):

    def __init__(self, bucket_bytes=25 << 20, backend=None, device=None):
        self.rank = int(os.environ.get("RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        self.backend = backend
        if backend == "nccl":
            torch.cuda.set_device(self.local_rank)
            self.device = torch.device("cuda:%d" % self.local_rank)
        else:
            self.device = torch.device(device or "cpu")
        if not dist.is_initialized():
            os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
            os.environ.setdefault("MASTER_PORT", "29500")
            kw = {}
            if backend == "nccl":
                kw["device_id"] = self.device
            dist.init_process_group(backend=backend, rank=self.rank, world_size=self.world, **kw)
        self.bucket_bytes = int(bucket_bytes)
        if os.environ.get("BASI_EXPERIMENTS") == "1" and os.environ.get("BASI_DP_BUCKET_MB"):
            self.bucket_bytes = int(float(os.environ["BASI_DP_BUCKET_MB"]) * (1 << 20))     # experiment: bucket size
        # (experiment BASI_DP_COMM_PRIORITY=1: the all-reduce stream at the highest priority, so the few NCCL CTAs win
        # the free SM slots at the kernel boundaries of the persistent backward kernels)
        prio = -1 if (os.environ.get("BASI_EXPERIMENTS") == "1" and os.environ.get("BASI_DP_COMM_PRIORITY") == "1") else 0
        self.comm_stream = torch.cuda.Stream(self.device, priority=prio) if backend == "nccl" else None
        self._plan = None
        # gradient exchange: "nccl" (default: bucketed NCCL all-reduce overlapped with the backward pass) or "p2p"
        # (BASI_DP_EXCHANGE=p2p: our peer-memory all-reduce after the backward pass; measured 0.36 ms of exposed time
        # at 2 GPUs against 0.21 ms for the overlapped NCCL buckets, so it is not the default)
        self.exchange = os.environ.get("BASI_DP_EXCHANGE", "nccl") if backend == "nccl" else "nccl"
        self.p2p = None
        self._p2p_tried = False

    # ---- simple (non-overlapped) form: callable on the flat gradient buffer
    def __call__(self, flat):
        return self.all_reduce_mean(flat)

    def all_reduce_mean(self, flat):
        if self.world == 1:
            return flat
        if self.backend == "nccl":
            dist.all_reduce(flat, op=dist.ReduceOp.AVG)
        else:
            dist.all_reduce(flat, op=dist.ReduceOp.SUM)
            flat.div_(self.world)
        return flat

    def broadcast(self, flat, src=0):
        if self.world > 1:
            dist.broadcast(flat, src)
        return flat

    def barrier(self):
        if self.world > 1:
            dist.barrier()

    # ---- bucket plan: which backward call completes which slice of the flat gradient buffer
    @staticmethod
    def plan_buckets(param_index, bwd_calls, n_flat, bucket_elems):
        """Returns [(call_index, start, end)]: after bwd_calls[call_index] has been enqueued, flat[start:end]
        is final.  Buckets are cut from the end of the buffer (the head's parameters finish first)."""
        last_writer = {}
        for i, call in enumerate(bwd_calls):
            for name in call[3].get("writes", ()):
                last_writer[name] = i
        names = list(param_index.keys())
        offs = [param_index[n][0] for n in names] + [n_flat]
        buckets = []
        end = n_flat
        ready = -1
        for j in range(len(names) - 1, -1, -1):
            ready = max(ready, last_writer.get(names[j], -1))
            start = offs[j]
            if end - start >= bucket_elems or j == 0:
                buckets.append((ready, start, end))
                end, ready = start, -1
        # a bucket must not be launched before an earlier-cut bucket (keeps every rank's collective order equal)
        out, hi = [], -1
        for ready, start, end in buckets:
            hi = max(hi, ready)
            out.append((hi, start, end))
        return out

    def run_backward(self, engine, st):
        """Runs engine.bwd on the current stream, launching bucket all-reduces on the side stream."""
        if self._plan is None:
            self._plan = self.plan_buckets(engine.param_index, engine.bwd, engine.n_flat,
                                           max(1, self.bucket_bytes // 4))
        calls = engine.bwd
        if self.exchange == "p2p" and self.world > 1 and not self._p2p_tried:
            self._p2p_tried = True
            try:
                self.p2p = P2PAllReduce(self.rank, self.world, engine.n_flat, self.device)
            except Exception as e:          # no symmetric memory on this system: the NCCL path below
                self.p2p = None
                if self.rank == 0:
                    sys.stderr.write("basi_b200.dp: peer-memory all-reduce unavailable (%s); using NCCL\n" % (e,))
            # every rank must take the same path
            ok = torch.tensor([1 if self.p2p is not None else 0], device=self.device)
            dist.all_reduce(ok, op=dist.ReduceOp.MIN)
            if int(ok.item()) == 0:
                self.p2p = None
        if self.p2p is not None:
            engine._run(calls, st)
            engine.join_side()
            self.p2p.all_reduce_mean(engine.grads_flat, st)
            return
        if self.comm_stream is None:
            # no side stream (gloo / CPU-side tests): backward, then one averaged all-reduce of the whole buffer
            engine._run(calls, st)
            engine.join_side()
            self.all_reduce_mean(engine.grads_flat)
            return
        cur = torch.cuda.current_stream(self.device)
        pos = 0
        for ready, start, end in self._plan:
            if ready + 1 > pos:
                engine._run(calls[pos:ready + 1], st)
                pos = ready + 1
            if self.world > 1:
                self.comm_stream.wait_stream(cur)
                engine.join_side(self.comm_stream)      # weight gradients run on the engine's side stream
                with torch.cuda.stream(self.comm_stream):
                    dist.all_reduce(engine.grads_flat[start:end], op=dist.ReduceOp.AVG)
        if pos < len(calls):
            engine._run(calls[pos:], st)
        if self.world > 1:
            cur.wait_stream(self.comm_stream)
