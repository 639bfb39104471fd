"""Lowers a ``Network`` graph (BAISPSPNet.py) to a static plan of C-ABI kernel calls and runs it.

This is the analogue of the reference's ``tf.Session``: the graph the fluent builder recorded is
compiled once into two flat call lists (forward, backward) over preallocated NHWC device buffers,
with the fusions the B200 design wants:

  * zero_padding + conv / atrous_conv            -> one implicit-GEMM conv with explicit pads
  * conv -> batch_normalization[+relu]           -> conv, bn_stats, bn_finalize, bn_apply(+ReLU)
  * add(shortcut, increase_bn) -> relu           -> one bn_apply with fused residual (and second BN) + ReLU
  * concat                                       -> producers write straight into channel slices
  * class_attention_conv (5x5/s5 on 5x5) and fc  -> skinny GEMMs
  * loss forward + gradient                      -> one kernel each

PyTorch is used for device memory, streams, CUDA graphs and NCCL only; every arithmetic op is one of
our kernels reached through ``_lib.call``.  There is no CPU path.
"""
from __future__ import annotations

import ctypes as C
import os
import re
from collections import OrderedDict

import numpy as np
import torch

from . import _lib
from ._lib import ConvDesc, Tensor

_ALIGN = 64  # parameter offsets aligned to 64 floats (256 B)
BN_REPLICAS = 8  # == BASI_BN_REPLICAS in include/basi_b200.h


def _same_pad_before(n_in, k, s, d):
    keff = (k - 1) * d + 1
    n_out = -(-n_in // s)
    total = max((n_out - 1) * s + keff - n_in, 0)
    return total // 2


def _exp_env(name):
    """Experiment switches count only together with BASI_EXPERIMENTS=1 -- a stray variable in a
    training job must not change results silently (same rule as basi::exp_env in the library)."""
    v = os.environ.get(name)
    if v is None:
        return None
    if os.environ.get("BASI_EXPERIMENTS") == "1":
        return v
    if name not in _WARNED:
        _WARNED.add(name)
        import sys
        sys.stderr.write("basi_b200: %s is set but ignored (experiment switches need BASI_EXPERIMENTS=1)\n" % name)
    return None


_WARNED = set()
_PSP_BRANCH = re.compile(r"(?:^|/)conv5_3_pool(\d)")     # pool1 / pool2 / pool3 / pool6 -> branch stream 0 / 1 / 2 / 3


class Act(object):
    """A device activation (NHWC torch tensor or channel-slice view) plus its C descriptor."""

    def __init__(self, t):
        assert t.dim() == 4 and t.stride(3) == 1
        n, h, w, c = t.shape
        ld = t.stride(2) if w > 1 else (t.stride(1) if h > 1 else (t.stride(0) if n > 1 else c))
        if w > 1 and h > 1:
            assert t.stride(1) == w * ld
        if h * w > 1 and n > 1:
            assert t.stride(0) == h * w * ld
        self.t = t
        self.dtype = _lib.F32 if t.dtype == torch.float32 else _lib.BF16
        self.desc = Tensor(t.data_ptr(), n, h, w, c, ld, self.dtype)
        self.ref = C.byref(self.desc)
        self.grad = None
        self.gw = False   # plan-time flag: has a backward op already written .grad ?

    @property
    def shape(self):
        return tuple(self.t.shape)


class _BNRec(object):
    __slots__ = ("x", "gamma", "beta", "sums", "dsums", "bnp", "coef", "count", "C", "cnt_f", "cnt_b")


class Engine(object):
    def __init__(self, net, batch_size, precision="bf16", training=True, loss=None, device=None, use_tc=True,
                 dry_run=False, fuse_bn_stats=True, overlap_wgrad=True, fuse_bn_bwd=True):
        """net: a built Network; loss: dict(kind='bce'|'softmax', pos_weight, class_weight, seg, cls).

        dry_run=True only builds the plan (buffers on the host, nothing can be executed): used by the
        CPU-side tests of the lowering."""
        self.dry_run = bool(dry_run)
        if self.dry_run:
            device, use_tc = "cpu", False
        elif not torch.cuda.is_available():
            raise _lib.BasiError("basi_b200 needs a CUDA device (sm_100a); there is no CPU path")
        if precision not in ("bf16", "f16", "f32"):
            raise ValueError("precision must be 'bf16', 'f16' or 'f32'")
        # precision "f16": the same kernels built for IEEE fp16 storage (libbasi_b200_f16.so).  3 more mantissa bits than
        # bf16 -- on this network the difference between missing and meeting the 2e-2 bar of a 16-bit path -- and a
        # narrower range: the loss gradient is scaled by `loss_scale` and the learning rate fed to SGD divided by it
        self.lib_fmt = "f16" if precision == "f16" else "bf16"
        _lib.use(self.lib_fmt)
        _lib.load()
        self.net = net
        # vgg_16 family (variant B, top-level LinkNet): split image / click-map inputs, biased slim convolutions,
        # per-map 2-channel losses.  (The cascade of back/8AttentionU also has `.attentions` but is a PSPNet.)
        self._vgg_family = hasattr(net, "attentions") and not getattr(net, "cascade", False)
        self.B = int(batch_size)
        self.precision = precision
        self.loss_scale = 1.0          # f16: set by _pick_loss_scale once the number of logits is known
        self.adt = torch.float32 if precision == "f32" else (torch.float16 if precision == "f16" else torch.bfloat16)
        # which tensors are stored in 16 bits (names of oracle.ROUNDING_POLICIES; tests hold the path to that model)
        self.storage_policy = "none" if precision == "f32" else "round1"
        self.training = training
        self.loss_cfg = loss
        if device is None:
            device = "cuda:%d" % torch.cuda.current_device()
        self.device = torch.device(device)
        self.use_tc = bool(use_tc)
        # f32 mode on the tensor cores: float32 storage, convolutions on bf16 [hi|mid|lo] split operands (six bf16
        # tcgen05 MMAs per product, fp32 TMEM accumulation) -- the reference's precision on the Blackwell-native path
        self.split_tc = self.use_tc and precision == "f32"
        # operand parts of the split mode: 3 = [hi|mid|lo], six products, float32-exact; 2 = [hi|mid], three products
        # (half the tensor-core work; 2^-17 unbiased operand error that averages out over the reduction)
        # Default: exact forward (3 parts: logits within 4e-6 of float64 at the benchmark shape) and three-product
        # backward (2 parts: its 4e-6 per-layer error is far below the 5e-3 the float32 reference itself loses on the
        # gradients there); BASI_SPLIT_PARTS=3 / 2 forces both passes
        sp = _exp_env("BASI_SPLIT_PARTS")
        self.split_parts = int(sp) if sp else 3
        self.split_parts_bwd = int(sp) if sp else 2
        self._split_cache = {}
        self._train_ranges = None
        self._bnact_of = {}
        # dgrad + BN backward in one cooperative launch (basi_tc_conv_set_bn_bwd): correct and tested, but measured
        # SLOWER than the separate resident BN-backward kernel (10.28 vs 9.09 ms/step: the two extra passes run on the
        # 8 epilogue warps of a one-CTA-per-SM kernel) -- opt-in experiment, BASI_FUSED_BWD=1
        self.fuse_bn_dgrad = not self.dry_run and precision != "f32" and _exp_env("BASI_FUSED_BWD") == "1"
        self.fused_bn_dgrad = 0
        self.fwd, self.bwd, self.pre = [], [], []
        self._keep = []          # ctypes objects that must outlive the plan
        self._ops = []
        self.overlap_wgrad = bool(overlap_wgrad) and not self.dry_run and not _exp_env("BASI_NO_OVERLAP")
        self._side = torch.cuda.Stream(self.device) if self.overlap_wgrad else None
        # weight-gradient scheduling: `wgrad_streams` side streams used round-robin -- the tensor-core weight-gradient
        # plans launch ~64 CTAs each, so two of them share the GPU (cfg3: 9.13 -> 8.81 ms/step).  `defer_wgrad`
        # (experiment, slower) moves every tensor-core weight gradient behind the dgrad / batch-norm chain
        ws = _exp_env("BASI_WGRAD_STREAMS")
        self.wgrad_streams = max(1, int(ws)) if ws else 2
        self.defer_wgrad = bool(_exp_env("BASI_DEFER_WGRAD")) and self.overlap_wgrad
        self._sides = ([self._side] + [torch.cuda.Stream(self.device) for _ in range(self.wgrad_streams - 1)]
                       if self.overlap_wgrad else [])
        self._side_rr = 0
        # SM partition of the backward pass (basi_set_sm_budget): the main chain (dgrad / batch-norm backward) sizes its
        # grids for `sm_main` SMs, every tensor-core weight gradient launches `sm_wgrad` CTAs, so the side stream owns
        # the remaining SMs instead of time-slicing with full-machine launches.  0 = off.
        self.sm_main = int(_exp_env("BASI_SM_MAIN") or 0) if self.overlap_wgrad else 0
        self.sm_wgrad = int(_exp_env("BASI_SM_WGRAD") or 0) if self.sm_main else 0
        # independent sub-graphs (the four PSP branches) can run on their own streams, forked/joined with events.
        # Measured inside the step graph: 10.27 ms with the branch streams vs 10.08 ms without (the extra cross-stream
        # dependencies cost more than the ~0.3 ms of tiny kernels they overlap), so this is opt-in.
        self._bstreams = None
        if not self.dry_run and _exp_env("BASI_BRANCH_STREAMS"):
            self._bstreams = [torch.cuda.Stream(self.device) for _ in range(4)]
        self._side_dirty = False
        self._tc_plans = []
        self._stem_producer = {}
        self._subsampled = {}
        self._tc_producer = {}
        self.fuse_bn_stats = fuse_bn_stats
        self.fused_stats = 0
        self.fuse_bn_apply = not self.dry_run and precision != "f32" and not _exp_env("BASI_NO_FUSED_APPLY")
        self.fused_apply = 0
        self.fuse_bn_bwd = bool(fuse_bn_bwd)
        self.mask_bits = not _exp_env("BASI_NO_MASK_BITS")
        # (the fused pyramid pooling accumulates with fp32 atomics: order noise of one ulp, which float32 storage keeps
        # and the batch-stat BN of the 1x1 pyramid branch -- statistics over only B samples -- amplifies; the f32
        # parity mode therefore uses the four separate, order-free pools)
        self.fuse_pools = not _exp_env("BASI_NO_POOL_FUSION") and precision != "f32"
        self.coop_bn_bwd = not _exp_env("BASI_NO_COOP_BN")
        self.fused_bn_bwd = 0
        self._tc_weights = []
        self._pack_table = None
        self._zero_grads = []    # gradient buffers that are pre-zeroed every step (concat buffers)
        self._acts = {}          # node index -> Act (or tuple for deferred bn)
        self._graph = None
        self.tc_layers = 0
        self._build_params()
        self._lower()
        if training:
            if self.sm_main:
                _lib.load().basi_set_sm_budget(self.sm_main, self.sm_wgrad)
            try:
                self._emit_backward()
            finally:
                if self.sm_main:
                    _lib.load().basi_set_sm_budget(0, 0)
            for c in self.bwd:
                c[3]["bwd"] = True
        if self.fused_apply >= 40:
            self.storage_policy = "fused"      # most BN inputs are never rounded: normalised from the accumulators

    # ------------------------------------------------------------------ parameters
    def _build_params(self):
        off = 0
        self.param_index = OrderedDict()
        for name, shape in self.net.variables.items():
            n = int(np.prod(shape))
            self.param_index[name] = (off, shape)
            off += -(-n // _ALIGN) * _ALIGN
        self.n_flat = off
        self.n_params = sum(int(np.prod(s)) for _, s in self.param_index.values())
        dev = self.device
        self.params_flat = self._zeros(off, torch.float32)
        self.grads_flat = self._zeros(off, torch.float32) if self.training else None
        self.lr_dev = self._zeros(1, torch.float32)

    def _pptr(self, name):
        return self.params_flat.data_ptr() + 4 * self.param_index[name][0]

    def _gptr(self, name):
        return self.grads_flat.data_ptr() + 4 * self.param_index[name][0]

    def param_view(self, name, grads=False):
        off, shape = self.param_index[name]
        flat = self.grads_flat if grads else self.params_flat
        return flat[off:off + int(np.prod(shape))].view(*shape)

    def set_params(self, params):
        """params: {tf variable name: ndarray} (conv weights HWIO)."""
        for name, (off, shape) in self.param_index.items():
            if name not in params:
                raise KeyError("missing variable %s" % name)
            v = np.ascontiguousarray(params[name], dtype=np.float32)
            if tuple(v.shape) != tuple(shape):
                raise ValueError("variable %s: shape %s != %s" % (name, v.shape, shape))
            self.param_view(name).copy_(torch.from_numpy(v))
        self._refresh_weight_copies()

    def set_trainable(self, substring=None):
        """A17: ``GradientDescentOptimizer.minimize(loss, var_list=[v for v in trainable if substring in v.name])`` --
        the class-only op of back/2AddClass/BAISRunnerTrain.py:120-121 ('class_attention') and the segment-side op of
        BAISRunnerTrain.py:54-58 ('segment_side').  None restores training of every variable.  Takes effect at the
        next step_device / capture."""
        if substring is None:
            self._train_ranges = None
            return 0
        ranges = []
        for name, (off, shape) in self.param_index.items():
            if substring in name:
                n = -(-int(np.prod(shape)) // _ALIGN) * _ALIGN
                if ranges and ranges[-1][0] + ranges[-1][1] == off:
                    ranges[-1][1] += n
                else:
                    ranges.append([off, n])
        if not ranges:
            raise KeyError("no variable name contains %r" % (substring,))
        self._train_ranges = [tuple(r) for r in ranges]
        self._graph = None                      # a captured step has the old optimizer launches
        return len(ranges)

    def broadcast_params(self, dp, src=0):
        """Data parallelism: rank `src`'s parameters to every replica, then refresh the bf16 tensor-core copies."""
        dp.broadcast(self.params_flat, src)
        self._refresh_weight_copies()

    def get_params(self):
        return OrderedDict((n, self.param_view(n).cpu().numpy().copy()) for n in self.param_index)

    def get_grads(self):
        """Parameter gradients of the last step (the loss scale of the f16 mode divided out)."""
        inv = 1.0 / self.loss_scale
        return OrderedDict((n, self.param_view(n, True).cpu().numpy() * np.float32(inv)) for n in self.param_index)

    def init_params(self, seed=0):
        """glorot_uniform for weights and biases (tf.get_variable default, BAISPSPNet.py:111-113), gamma=1, beta=0."""
        rng = np.random.RandomState(seed)
        p = {}
        for name, (off, shape) in self.param_index.items():
            if name.endswith("/gamma"):
                v = np.ones(shape)
            elif name.endswith("/beta"):
                v = np.zeros(shape)
            else:
                if len(shape) == 4:
                    fi, fo = shape[0] * shape[1] * shape[2], shape[0] * shape[1] * shape[3]
                elif len(shape) == 2:
                    fi, fo = shape
                else:
                    fi = fo = shape[0]
                lim = np.sqrt(6.0 / (fi + fo))
                v = rng.uniform(-lim, lim, shape)
            p[name] = v.astype(np.float32)
        self.set_params(p)

    # ------------------------------------------------------------------ allocation helpers
    def _zeros(self, shape, dtype):
        """Zero-initialised device buffer: torch.empty + basi_memset (a stream memset, not a fill kernel -- the plan
        has ~700 buffers and a fill kernel per buffer would bury our kernels in a profiler's launch list)."""
        if isinstance(shape, int):
            shape = (shape,)
        if self.dry_run or self.device.type != "cuda":
            return torch.zeros(*shape, dtype=dtype, device=self.device)
        t = torch.empty(*shape, dtype=dtype, device=self.device)
        if t.numel():
            _lib.call("basi_memset", t.data_ptr(), 0, C.c_int64(t.numel() * t.element_size()),
                      torch.cuda.current_stream(self.device).cuda_stream)
        return t

    def _alloc(self, shape, dtype=None):
        return self._zeros(tuple(shape), dtype or self.adt)

    def _new_act(self, h, w, c, dtype=None):
        return Act(self._alloc((self.B, h, w, c), dtype))

    def _grad_of(self, act):
        if act.grad is None:
            act.grad = Act(self._zeros(tuple(act.t.shape), act.t.dtype))
        return act.grad

    def _call(self, lst, name, *args, **meta):
        """Appends one C-ABI call; meta: flops / bytes (algorithmic, for the roofline) and writes=[param names]."""
        fn = getattr(_lib.load(), name)
        if getattr(self, "_cur_branch", None) is not None:
            meta["branch"] = self._cur_branch
        lst.append((name, fn, args, meta))

    # ------------------------------------------------------------------ lowering
    def _lower(self):
        net = self.net
        B = self.B
        nodes = net.nodes
        cons = {n.index: [] for n in nodes}
        for n in nodes:
            for i in n.inputs:
                cons[i.index].append(n)
        self._cons = cons
        # BN statistic scratch: [fwd sums | bwd sums] as doubles, zeroed once per step
        tot_c = sum(n.shape[-1] for n in nodes if n.op == "batch_normalization")
        n_bn = sum(1 for n in nodes if n.op == "batch_normalization")
        self.bn_scratch = self._zeros(4 * tot_c * BN_REPLICAS + 8, torch.float64)
        self.bn_counters = self._zeros(2 * n_bn + 2, torch.int32)   # last-block tickets
        self._bn_off = 0
        self._bn_idx = 0
        # concat: producers write into slices
        self._slice_of = {}
        for n in nodes:
            if n.op == "concat":
                h, w, ctot = n.shape
                buf = self._new_act(h, w, ctot, torch.float32 if n.attrs.get("dtype") == "f32" else None)
                self._acts[n.index] = buf
                o = 0
                for i in n.inputs:
                    c = i.shape[2]
                    self._slice_of[i.index] = (buf, o, c)
                    o += c
        self.input = None
        self.seg_logits = None
        self.cls_logits = None
        for n in nodes:
            # the four pyramid-pooling branches (pool -> 1x1 conv -> BN -> bilinear into a concat slice) are independent
            # chains of tiny kernels: their calls are tagged and run on per-branch streams (see _run)
            m = _PSP_BRANCH.search(n.name or "")
            self._cur_branch = {1: 0, 2: 1, 3: 2, 6: 3}.get(int(m.group(1)), int(m.group(1)) % 4) if m else None
            first_op = len(self._ops)
            getattr(self, "_lower_" + n.op)(n)
            for _, op in self._ops[first_op:]:
                op["branch"] = self._cur_branch
        self._cur_branch = None
        self._lower_loss()

    def _out_act(self, node, dtype=None):
        """Output buffer for a node: a concat slice when the node feeds a concat, else a fresh tensor."""
        if node.index in self._slice_of:
            buf, o, c = self._slice_of[node.index]
            a = Act(buf.t[..., o:o + c])
            if self.training:
                if buf.grad is None:
                    self._zero_grads.append(self._grad_of(buf))
                a.grad = Act(buf.grad.t[..., o:o + c])
                a.gw = True            # concat grads are pre-zeroed -> every writer accumulates
                buf.gw = True
            return a
        h, w, c = node.shape
        return self._new_act(h, w, c, dtype)

    def _lower_data(self, n):
        h, w, c = n.shape
        if self._vgg_family:
            # variant B: image (3 ch) and click map (1 ch) are two placeholders; both are channel-slice views of ONE
            # NHWC4 float32 buffer so that basi_clickmap_pack (uint8 image + click -> NHWC4) feeds them directly
            if getattr(self, "_input4", None) is None:
                self._input4 = self._alloc((self.B, h, w, 4), torch.float32)
            a = Act(self._input4[..., 3:4] if n.name == "mask" else self._input4[..., 0:3])
        else:
            a = Act(self._alloc((self.B, h, w, c), torch.float32))
        self._acts[n.index] = a
        if n.name == "mask":              # the click map needs no gradient
            self.input_mask = a
            a.no_grad = True
        else:
            self.input = a

    def _lower_zero_padding(self, n):
        self._acts[n.index] = ("pad", self._acts[n.inputs[0].index], n.attrs["pad"])

    def _lower_conv(self, n):
        a = n.attrs
        src = self._acts[n.inputs[0].index]
        pad = 0
        if isinstance(src, tuple) and src[0] == "pad":
            _, src, pad = src
        x = src
        k, s, d = a["k_h"], a["stride"], a["dilation"]
        if k == 1 and a["k_w"] == 1 and s > 1 and a["padding"] != "SAME" and pad == 0:
            # strided 1x1 VALID conv == stride-1 1x1 conv of x[:, ::s, ::s, :]; the subsampled tensor is shared by all
            # strided convs reading the same input (conv3_1 proj and reduce)
            key = (id(x), s)
            if key not in self._subsampled:
                xs = self._new_act(-(-x.shape[1] // s), -(-x.shape[2] // s), x.shape[3],
                                   torch.float32 if x.t.dtype == torch.float32 else None)
                self._subsampled[key] = xs
                sop = dict(x=x, y=xs, stride=s)
                self._ops.append(("subsample", sop))
                self._call(self.fwd, "basi_subsample_fwd", x.ref, s, xs.ref)
            x = self._subsampled[key]
            s = 1
        if a["padding"] == "SAME":
            pt, pl = _same_pad_before(x.shape[1], k, s, d), _same_pad_before(x.shape[2], a["k_w"], s, d)
        else:
            pt = pl = pad
        oh, ow, co = n.shape
        followers = self._cons[n.index]
        has_bn = any(f.op == "batch_normalization" for f in followers)
        wname, bname = a["weights"], a["biases"]
        # class_attention_conv pattern: 5x5/s5 over a dense 5x5 map -> skinny GEMM
        if (k == 5 and s == 5 and x.shape[1] == 5 and x.shape[2] == 5 and pt == 0 and oh == 1 and ow == 1
                and x.desc.ld == x.shape[3]
                and _lib.load().basi_skinny_supported(self.B, 25 * x.shape[3], co) == 1):
            y = Act(self._alloc((self.B, 1, 1, co), torch.float32))
            self._acts[n.index] = y
            self._ops.append(("skinny", dict(x=x, y=y, w=wname, b=bname, relu=bool(a["relu"]),
                                             K=25 * x.shape[3], N=co)))
            self._emit_skinny_fwd(self._ops[-1][1])
            return
        # slim conv2d of the vgg_16 trunk (variant B): bias + ReLU, no BN.  In bf16 mode these stay bf16 so that they
        # run on the tcgen05 path (bias + ReLU in its epilogue); every other BN-less conv is a float32 head
        vgg_like = (self._vgg_family and self.precision != "f32" and not has_bn and s == 1 and co > 4)
        f32_out = (not has_bn) and not vgg_like
        y = self._out_act(n, torch.float32 if f32_out else None)
        desc = ConvDesc(k, a["k_w"], s, d, pt, pl, 1 if (a["relu"] and not has_bn) else 0)
        self._keep.append(desc)
        op = dict(x=x, y=y, w=wname, b=bname, desc=desc, relu=bool(a["relu"] and not has_bn), name=n.name, tc=None)
        self._acts[n.index] = y
        self._ops.append(("conv", op))
        self._emit_conv_fwd(op)

    def _lower_batch_normalization(self, n):
        x = self._acts[n.inputs[0].index]
        rec = _BNRec()
        rec.x, rec.gamma, rec.beta, rec.C = x, n.attrs["gamma"], n.attrs["beta"], n.shape[-1]
        rec.count = float(x.shape[0] * x.shape[1] * x.shape[2])
        Cc = rec.C
        base = self.bn_scratch.data_ptr()
        rec.sums = base + 8 * self._bn_off
        rec.dsums = base + 8 * (self._bn_off + 2 * Cc * BN_REPLICAS)
        self._bn_off += 4 * Cc * BN_REPLICAS
        rec.cnt_f = self.bn_counters.data_ptr() + 8 * self._bn_idx
        rec.cnt_b = rec.cnt_f + 4
        self._bn_idx += 1
        rec.bnp = self._alloc((4 * Cc,), torch.float32)
        rec.coef = self._alloc((2 * Cc,), torch.float32)
        followers = self._cons[n.index]
        if followers and all(f.op == "add" for f in followers):
            self._acts[n.index] = ("bn_deferred", rec)     # applied inside the junction
            self._emit_bn_stats(rec)
            return
        relu = bool(n.attrs["relu"])
        if len(followers) == 1 and followers[0].op == "relu":
            relu = True                                      # bn(relu=False) -> relu (conv1_1)
            out = self._out_act(followers[0])
            self._acts[followers[0].index] = out
        else:
            out = self._out_act(n)
        if not (len(followers) == 1 and followers[0].op == "relu"):
            self._acts[n.index] = out
        else:
            self._acts[n.index] = ("fused_into_relu", out)
        self._emit_bn_stats(rec)
        op = dict(main=rec, res=None, res_bn=None, relu=relu, out=out)
        self._ops.append(("bnact", op))
        outnode = followers[0] if (len(followers) == 1 and followers[0].op == "relu") else n
        self._bnact_of[id(out)] = (op, len(self._cons[outnode.index]))     # for the fused BN backward (see _bwd_conv)
        prod = self._tc_producer.get(id(rec.x))
        if (prod is not None and self.fuse_bn_stats and self.fuse_bn_apply and not prod.get("split")
                and _lib.load().basi_tc_conv_set_bn_apply(prod["tc_fprop"], out.ref, 1 if relu else 0) == 1):
            # conv -> statistics -> grid barrier -> normalise + ReLU from the fp32 TMEM accumulators, one launch
            op["fused_apply"] = True
            self.fused_apply += 1
            return
        self._emit_bnact_fwd(op)

    def _lower_relu(self, n):
        src = n.inputs[0]
        if src.op == "batch_normalization":
            assert n.index in self._acts, "relu after bn must have been fused"
            return
        if src.op != "add":
            raise NotImplementedError("relu after %s" % src.op)
        if isinstance(self._acts.get(src.index), Act):
            # the sum itself was materialised (it has readers of its own, see _lower_add): a stand-alone ReLU
            x = self._acts[src.index]
            y = self._out_act(n, x.t.dtype)
            self._acts[n.index] = y
            self._ops.append(("relu", dict(x=x, y=y)))
            self._call(self.fwd, "basi_relu_fwd", x.ref, y.ref, bytes=self._nbytes(x) * 2)
            return
        main, res, res_bn = self._junction_inputs(src)
        out = self._out_act(n)
        self._acts[n.index] = out
        self._acts[src.index] = ("pre_relu_of", out)
        op = dict(main=main, res=res, res_bn=res_bn, relu=True, out=out)
        self._ops.append(("bnact", op))
        self._emit_bnact_fwd(op)

    def _junction_inputs(self, add_node):
        """(main BN record, residual tensor, residual BN record or None) of a residual junction."""
        ins = [self._acts[i.index] for i in add_node.inputs]
        assert len(ins) == 2
        deferred = [i for i in ins if isinstance(i, tuple) and i[0] == "bn_deferred"]
        plain = [i for i in ins if isinstance(i, Act)]
        if len(deferred) == 2:
            # (proj_bn, increase_bn): main = the later one
            res_bn, main = deferred[0][1], deferred[1][1]
            return main, res_bn.x, res_bn
        if len(deferred) == 1 and len(plain) == 1:
            return deferred[0][1], plain[0], None
        raise NotImplementedError("add of %s" % (ins,))

    def _lower_add(self, n):
        ins = [self._acts[i.index] for i in n.inputs]
        if len(ins) == 2 and all(isinstance(i, Act) for i in ins):
            # plain tensor add without batch norm (top-level LinkNet, BAISNet.py:244)
            a, b = ins
            y = self._out_act(n, a.t.dtype)
            self._acts[n.index] = y
            self._ops.append(("add", dict(a=a, b=b, y=y)))
            self._call(self.fwd, "basi_add_fwd", a.ref, b.ref, y.ref)
            return
        if any(c.op not in ("relu", "multiply") for c in self._cons[n.index]):      # (multiply has its ReLU inside)
            # a residual junction whose SUM is read besides its ReLU (the hand-unrolled trunk of
            # back/8AttentionU/BAISNet.py:163-165 feeds it to the next 1x1_reduce / 1x1_proj): materialise
            # BN(main) + residual without the ReLU; the ReLU that follows becomes its own op (_lower_relu).
            # The adjoint runs on the generic reduce / apply pair (no stored-output mask, residual gradient).
            main, res, res_bn = self._junction_inputs(n)
            out = self._out_act(n)
            self._acts[n.index] = out
            op = dict(main=main, res=res, res_bn=res_bn, relu=False, out=out, generic_bwd=True)
            self._ops.append(("bnact", op))
            self._emit_bnact_fwd(op)
        # else: a residual junction, materialised by the following relu

    def _lower_concat(self, n):
        pass  # producers already wrote into the slices

    def _lower_max_pool(self, n):
        x = self._acts[n.inputs[0].index]
        if n.attrs.get("k", 3) == 2:       # slim vgg_16 pools (variant B)
            y = self._out_act(n, x.t.dtype)
            amax = self._zeros(self.B * n.shape[0] * n.shape[1] * n.shape[2], torch.uint8)
            self._acts[n.index] = y
            self._ops.append(("maxpool2", dict(x=x, y=y, amax=amax)))
            self._call(self.fwd, "basi_maxpool2s2_fwd", x.ref, y.ref, amax.data_ptr())
            return
        y = self._out_act(n)
        amax = self._zeros(self.B * n.shape[0] * n.shape[1] * n.shape[2], torch.uint8)
        op = dict(x=x, y=y, amax=amax)
        self._acts[n.index] = y
        self._ops.append(("maxpool", op))
        self._call(self.fwd, "basi_maxpool3s2_fwd", x.ref, y.ref, amax.data_ptr())

    def _lower_avg_pool(self, n):
        if n.index in self._acts:
            return                                   # already produced by the fused pass of its pyramid group
        x = self._acts[n.inputs[0].index]
        # pyramid pooling: every avg_pool reading this tensor is computed in ONE pass, emitted at the first of them
        sibs = [c for c in self._cons[n.inputs[0].index] if c.op == "avg_pool"]
        if len(sibs) > 4:      # the cascade (F4): a gated feature feeds a class-head pool and the next decoder's pyramid
            sibs = [c for c in sibs if _PSP_BRANCH.search(c.name or "")]
        if self.fuse_pools and 2 <= len(sibs) <= 4 and n in sibs and self._pool_group_fits(x, sibs):
            grp = dict(x=x, ops=[])
            for sib in sibs:
                op = dict(x=x, y=self._out_act(sib), k=sib.attrs["k"], group=grp)
                self._acts[sib.index] = op["y"]
                self._ops.append(("avgpool", op))
                grp["ops"].append(op)
            self._emit_pool_group_fwd(grp)
            return
        y = self._out_act(n)
        op = dict(x=x, y=y, k=n.attrs["k"], group=None)
        self._acts[n.index] = y
        self._ops.append(("avgpool", op))
        self._call(self.fwd, "basi_avgpool_fwd", x.ref, op["k"], y.ref)

    def _pool_group_fits(self, x, sibs):
        ks = (C.c_int * len(sibs))(*[sb.attrs["k"] for sb in sibs])
        return _lib.load().basi_avgpool_multi_scratch_floats(C.byref(x.desc), len(sibs), ks) > 0

    def _pool_group_args(self, grp, grads):
        ops = grp["ops"]
        ks = (C.c_int * len(ops))(*[o["k"] for o in ops])
        ptrs = (C.POINTER(Tensor) * len(ops))(*[C.pointer((o["y"].grad if grads else o["y"]).desc) for o in ops])
        self._keep.extend([ks, ptrs])
        return len(ops), ks, ptrs

    def _emit_pool_group_fwd(self, grp):
        x = grp["x"]
        n, ks, ptrs = self._pool_group_args(grp, False)
        nfl = int(_lib.load().basi_avgpool_multi_scratch_floats(x.ref, n, ks))    # per-row window sums [n][h][cells][c]
        grp["scratch"] = self._zeros(nfl, torch.float32)
        br, self._cur_branch = self._cur_branch, None          # the shared pass runs before the branches fork
        self._call(self.fwd, "basi_avgpool_multi_fwd", x.ref, n, ks, ptrs, grp["scratch"].data_ptr(),
                   bytes=self._nbytes(x))
        self._cur_branch = br

    def _lower_resize_bilinear(self, n):
        x = self._acts[n.inputs[0].index]
        y = self._out_act(n)
        op = dict(x=x, y=y)
        self._acts[n.index] = y
        self._ops.append(("bilinear", op))
        self._call(self.fwd, "basi_bilinear_ac_fwd", x.ref, y.ref)

    def _lower_multiply(self, n):
        feat = self._acts[n.inputs[0].index]
        if isinstance(feat, tuple) and feat[0] == "pre_relu_of":
            feat = feat[1]      # relu(pre) == the junction output
        logits = self._acts[n.inputs[1].index]
        assert logits.t.dtype == torch.float32
        y = self._out_act(n)
        nseg = logits.shape[3]
        att = n.attrs["segment_place"] if nseg > 1 else 0
        op = dict(feat=feat, logits=logits, y=y, nseg=nseg, att=att)
        self._acts[n.index] = y
        self._ops.append(("gate", op))
        self._call(self.fwd, "basi_gate_mul_fwd", feat.ref, logits.t.data_ptr(), nseg, att, y.ref)

    # ---- variant B (F1): nearest resize, click / attention gating, softmax gate
    def _mask_grad(self, act):
        """1-channel float32 maps (attention gates and their resized copies) collect gradient from several consumers
        through kernels that always ADD (basi_mask_mul_bwd): the buffer is zeroed every step."""
        if getattr(act, "no_grad", False):
            return None
        if act.grad is None:
            self._zero_grads.append(self._grad_of(act))
            act.gw = True
        return act.grad

    def _lower_resize_nearest(self, n):
        x = self._acts[n.inputs[0].index]
        y = self._out_act(n, x.t.dtype)
        if getattr(x, "no_grad", False):
            y.no_grad = True
        self._acts[n.index] = y
        self._ops.append(("nearest", dict(x=x, y=y)))
        self._call(self.fwd, "basi_resize_nearest_fwd", x.ref, y.ref)

    def _lower_mask_multiply(self, n):
        feat = self._acts[n.inputs[0].index]
        mask = self._acts[n.inputs[1].index]
        assert mask.t.dtype == torch.float32 and mask.shape[3] == 1
        y = self._out_act(n, feat.t.dtype)
        self._acts[n.index] = y
        self._ops.append(("maskmul", dict(feat=feat, mask=mask, y=y)))
        self._call(self.fwd, "basi_mask_mul_fwd", feat.ref, mask.t.data_ptr(), 1, 0, y.ref)

    def _lower_softmax_gate(self, n):
        logits = self._acts[n.inputs[0].index]
        assert logits.t.dtype == torch.float32
        h, w, c = n.inputs[0].shape
        gate = Act(self._alloc((self.B, h, w, 1), torch.float32))
        self._acts[n.index] = gate
        op = dict(logits=logits, gate=gate, C=c, sel=int(n.attrs["sel"]), thr=float(n.attrs["thr"]))
        self._ops.append(("softgate", op))
        self._call(self.fwd, "basi_softmax_gate_fwd", logits.t.data_ptr(), C.c_int64(self.B * h * w), c, op["sel"],
                   C.c_float(op["thr"]), gate.t.data_ptr())

    def _lower_sigmoid(self, n):
        x = self._acts[n.inputs[0].index]
        assert x.t.dtype == torch.float32
        h, w, c = n.shape
        y = Act(self._alloc((self.B, h, w, c), torch.float32))
        self._acts[n.index] = y
        self._ops.append(("sigmoid", dict(x=x, y=y, n=self.B * h * w * c)))
        self._call(self.fwd, "basi_sigmoid_fwd", x.t.data_ptr(), y.t.data_ptr(), C.c_int64(self.B * h * w * c))

    def _lower_squeeze(self, n):
        self._acts[n.index] = self._acts[n.inputs[0].index]

    def _lower_fc(self, n):
        x = self._acts[n.inputs[0].index]
        assert x.t.dtype == torch.float32 and x.shape[1] == 1 and x.shape[2] == 1
        co = n.shape[0]
        if _lib.load().basi_skinny_supported(self.B, x.shape[3], co) != 1:
            raise _lib.BasiError("fc %s: shape (B=%d, K=%d, N=%d) not supported by the skinny GEMM"
                                 % (n.name, self.B, x.shape[3], co))      # fails at plan time, never mid-step
        y = Act(self._alloc((self.B, 1, 1, co), torch.float32))
        self._acts[n.index] = y
        op = dict(x=x, y=y, w=n.attrs["weights"], b=n.attrs["biases"], relu=bool(n.attrs["relu"]),
                  K=x.shape[3], N=co)
        self._ops.append(("skinny", op))
        self._emit_skinny_fwd(op)

    @staticmethod
    def _conv_flops(op):
        y, d, x = op["y"], op["desc"], op["x"]
        return 2.0 * y.shape[0] * y.shape[1] * y.shape[2] * y.shape[3] * d.kh * d.kw * x.shape[3]

    # ---- forward emitters
    def _emit_conv_fwd(self, op):
        x, y = op["x"], op["y"]
        bptr = self._pptr(op["b"]) if op["b"] else None
        if self.use_tc and x.dtype == _lib.BF16 and y.dtype == _lib.BF16 and (not op["b"] or op["relu"]):
            # (the descriptor's relu flag belongs to the CUDA-core path; the tcgen05 plan takes bias / ReLU separately)
            d0 = ConvDesc(op["desc"].kh, op["desc"].kw, op["desc"].stride, op["desc"].dil, op["desc"].pad_t,
                          op["desc"].pad_l, 0)
            if _lib.load().basi_tc_conv_supported(_lib.TC_FPROP, C.byref(d0), x.ref, y.ref) == 1:
                if op["b"]:
                    self._keep.append(d0)
                    op["desc_simt"], op["desc"], op["tc_bias"] = op["desc"], d0, True
                op["tc_fprop"] = self._emit_tc(op, _lib.TC_FPROP, self.fwd)
                if op["b"]:
                    _lib.call("basi_tc_conv_set_bias", op["tc_fprop"], bptr, 1 if op["relu"] else 0)
                else:
                    self._tc_producer[id(y)] = op
                return
        if self.split_tc and x.dtype == _lib.F32 and y.dtype == _lib.F32 and not op["b"] and not self.dry_run:
            if _lib.load().basi_tc_conv_supported_split(_lib.TC_FPROP, C.byref(op["desc"]), x.ref, y.ref) == 1:
                op["split"] = True
                op["tc_fprop"] = self._emit_tc(op, _lib.TC_FPROP, self.fwd)
                self._tc_producer[id(y)] = op
                return
        if (not op["b"] and not self.dry_run and self.fuse_bn_stats
                and _lib.load().basi_stem_fprop_stats_supported(C.byref(op["desc"]), x.ref, y.ref) == 1):
            self._stem_producer[id(y)] = (len(self.fwd), op)       # _emit_bn_stats may upgrade this call
        self._call(self.fwd, "basi_conv_fprop", C.byref(op["desc"]), x.ref, self._pptr(op["w"]), bptr, y.ref,
                   flops=self._conv_flops(op))

    @staticmethod
    def _nbytes(act):
        return float(np.prod(act.shape)) * act.t.element_size()

    def _emit_bn_stats(self, rec):
        prod = self._tc_producer.get(id(rec.x))
        if prod is not None and self.fuse_bn_stats:
            # statistics come out of the tcgen05 conv epilogue: no separate pass over the conv output
            _lib.call("basi_tc_conv_set_bn_stats", prod["tc_fprop"], rec.sums, self._pptr(rec.gamma),
                      self._pptr(rec.beta), C.c_double(rec.count), C.c_float(1e-5), rec.bnp.data_ptr(), rec.cnt_f)
            self.fused_stats += 1
            return
        stem = self._stem_producer.get(id(rec.x))
        if stem is not None:
            # conv1_1: the stem kernel accumulates the statistics of its own output
            pos, op = stem
            name, _, args, meta = self.fwd[pos]
            assert name == "basi_conv_fprop"
            fn = _lib.load().basi_stem_fprop_stats
            self.fwd[pos] = ("basi_stem_fprop_stats", fn,
                             (args[0], args[1], args[2], args[4], rec.sums, self._pptr(rec.gamma), self._pptr(rec.beta),
                              C.c_double(rec.count), C.c_float(1e-5), rec.bnp.data_ptr(), rec.cnt_f), meta)
            self.fused_stats += 1
            return
        self._call(self.fwd, "basi_bn_stats", rec.x.ref, rec.sums, self._pptr(rec.gamma), self._pptr(rec.beta),
                   C.c_double(rec.count), C.c_float(1e-5), rec.bnp.data_ptr(), rec.cnt_f, bytes=self._nbytes(rec.x))

    def _emit_bnact_fwd(self, op):
        main, res, res_bn = op["main"], op["res"], op["res_bn"]
        op["bits"] = None
        if (res is not None and op["relu"] and self.training and self.mask_bits and not self.dry_run
                and _lib.load().basi_bn_maskbits_supported(main.x.ref) == 1):
            # residual junction: also emit the packed ReLU mask, the backward pair reads it instead of `out`
            n_, h_, w_, c_ = main.x.shape
            op["bits"] = self._zeros(n_ * h_ * w_ * (c_ // 8), torch.uint8)
            self._call(self.fwd, "basi_bn_apply_bits", main.x.ref, main.bnp.data_ptr(), res.ref,
                       res_bn.bnp.data_ptr() if res_bn is not None else None, 1, op["out"].ref,
                       op["bits"].data_ptr(), bytes=self._nbytes(main.x) * (3 + 1.0 / 16))
            return
        self._call(self.fwd, "basi_bn_apply", main.x.ref, main.bnp.data_ptr(), res.ref if res is not None else None,
                   res_bn.bnp.data_ptr() if res_bn is not None else None, 1 if op["relu"] else 0, op["out"].ref,
                   bytes=self._nbytes(main.x) * (2 + (1 if res is not None else 0)))

    def _emit_skinny_fwd(self, op):
        x, y = op["x"], op["y"]
        lda = x.t.stride(0) if x.shape[0] > 1 else op["K"]
        op["lda"] = lda
        # the workspace variant adds the k-split partial sums in a fixed order: with the atomic variant the class logits
        # differ by 1e-7 from run to run, which the 16-bit backward amplifies to 1.5e-2 on the weight gradients
        nws = int(_lib.load().basi_skinny_fwd_workspace_floats(self.B, op["K"], op["N"]))
        op["ws"] = self._zeros((max(nws, 1),), torch.float32)
        self._call(self.fwd, "basi_skinny_fwd_ws", x.t.data_ptr(), x.dtype, C.c_int64(lda), self._pptr(op["w"]),
                   self._pptr(op["b"]) if op["b"] else None, y.t.data_ptr(), self.B, op["K"], op["N"],
                   1 if op["relu"] else 0, op["ws"].data_ptr())

    # ---- loss
    def _pick_loss_scale(self, n_logits):
        """f16 mode: static loss scale, a power of two that puts the scaled loss gradient of one logit (scale / N) near
        0.25 -- 4096 at the benchmark shape (N = 25 600), 32 for a 64 x 64 toy batch -- so that neither the small
        activation gradients underflow fp16 nor the batch-norm adjoints (x gamma * istd) overflow it."""
        if self.precision != "f16":
            return
        s = 1.0
        while s * 2 <= min(8192.0, n_logits / 4.0):
            s *= 2
        self.loss_scale = s

    def _lower_loss_linknet(self):
        """cal_loss of variant B (back/90AttentionSingle2/BAISRunnerTrain.py:116-156): every attention map against the
        nearest-resized labels as 2-channel weighted CE (pos_weight 3, mean over 2N elements), averaged over the maps,
        plus the class-head softmax CE (weight 1)."""
        cfg = self.loss_cfg or {}
        net = self.net
        heads = [self._acts[nd.index] for nd in net.attentions]
        self.att_logits = heads
        self.seg_logits = heads[-1]
        self.seg_name = "attention_1"
        self.cls_logits = self._acts[net.classes[0].index] if net.classes else None
        self.cls_name = "class_attention_fc" if net.classes else None
        B = self.B
        S_h, S_w = self.input.shape[1], self.input.shape[2]
        stride = int(getattr(net, "label_stride", 8))          # variant B: labels at S/8; top-level LinkNet: full size
        P_h, P_w = S_h // stride, S_w // stride
        self.loss_acc = self._zeros(4, torch.float64)
        self.pred_seg = self._zeros((B, heads[-1].shape[1], heads[-1].shape[2], 1), torch.int32)
        self.pred_cls = self._zeros((B,), torch.int32) if self.cls_logits is not None else None
        self.post = []
        fin = heads[-1]
        self._call(self.post, "basi_argmax", fin.t.data_ptr(), C.c_int64(B * fin.shape[1] * fin.shape[2]), 2,
                   self.pred_seg.data_ptr())
        if self.cls_logits is not None:
            self._call(self.post, "basi_argmax", self.cls_logits.t.data_ptr(), C.c_int64(B),
                       self.cls_logits.shape[3], self.pred_cls.data_ptr())
        if not self.training:
            return
        self.label_seg = self._zeros((B, P_h, P_w, 1), torch.float32)
        lab = Act(self.label_seg)
        self._pick_loss_scale(B * heads[0].shape[1] * heads[0].shape[2] * 2)
        self.lossl = []
        self._lab_scaled = []
        pw = float(cfg.get("pos_weight", 3.0))
        for a in heads:
            _, h, w, c = a.shape
            assert c == 2
            li = Act(self._zeros((B, h, w, 1), torch.float32))
            ti = self._zeros((B, h, w, 2), torch.float32)
            self._lab_scaled.append((li, ti))
            n2 = 2 * B * h * w
            g = self._grad_of(a)
            a.gw = True
            self._call(self.lossl, "basi_resize_nearest_fwd", lab.ref, li.ref)
            self._call(self.lossl, "basi_onehot2_f32", li.t.data_ptr(), ti.data_ptr(), C.c_int64(B * h * w))
            self._call(self.lossl, "basi_wbce_fwd_bwd", a.t.data_ptr(), ti.data_ptr(), C.c_float(pw),
                       C.c_double(1.0 / (len(heads) * n2)), C.c_float(self.loss_scale / (len(heads) * n2)), C.c_int64(n2),
                       self.loss_acc.data_ptr(), g.t.data_ptr())
        self.class_weight = float(cfg.get("class_weight", 1.0))
        if self.cls_logits is not None:
            self.label_cls = self._zeros((B,), torch.int32)
            gc = self._grad_of(self.cls_logits)
            self.cls_logits.gw = True
            ncls = self.cls_logits.shape[3]
            self._call(self.lossl, "basi_softmax_ce_fwd_bwd", self.cls_logits.t.data_ptr(),
                       self.label_cls.data_ptr(), C.c_int64(B), ncls, C.c_double(1.0 / B),
                       C.c_float(self.loss_scale * self.class_weight / B), self.loss_acc.data_ptr() + 8, gc.t.data_ptr())
        else:
            self.label_cls = None

    def _lower_loss_cascade(self):
        """cal_loss of the cascade (back/8AttentionU/BAISRunnerTrain.py:161-193) on the SIGMOID outputs: softmax CE
        against the 4-class labels for the first decoders, 2 x weighted BCE (pos_weight 3) of channel 1 against the
        attention labels for the last `attention_module_num`; loss = mean(segment terms) + 0.1 * mean(class CEs).
        Predictions come from segments[0] / classes[0] (:63-69)."""
        cfg = self.loss_cfg or {}
        net = self.net
        segs = [self._acts[nd.index] for nd in net.segments]
        clss = [self._acts[nd.index] for nd in net.classes]
        self.segments, self.classes_logits = segs, clss
        self.seg_logits, self.seg_name = segs[0], net.segments[0].name
        self.cls_logits, self.cls_name = clss[0], net.classes[0].name
        B, P_h, P_w, nseg = segs[0].shape
        N = B * P_h * P_w
        self.loss_acc = self._zeros(4, torch.float64)
        self.pred_seg = self._zeros((B, P_h, P_w, 1), torch.int32)
        self.pred_cls = self._zeros((B,), torch.int32)
        self.post = []
        self._call(self.post, "basi_argmax", segs[0].t.data_ptr(), C.c_int64(N), nseg, self.pred_seg.data_ptr())
        self._call(self.post, "basi_argmax", clss[0].t.data_ptr(), C.c_int64(B), clss[0].shape[3],
                   self.pred_cls.data_ptr())
        if not self.training:
            return
        self._pick_loss_scale(N)
        self.label_seg = self._zeros((B, P_h, P_w, 1), torch.int32)
        self.label_att = self._zeros((B, P_h, P_w, 1), torch.float32)
        self.label_cls = self._zeros((B,), torch.int32)
        self.lossl = []
        pw = float(cfg.get("pos_weight", 3.0))
        n_ce = len(segs) - int(net.attention_module_num)
        for i, a in enumerate(segs):
            g = self._grad_of(a)
            a.gw = True
            c = a.shape[3]
            if i < n_ce:
                self._call(self.lossl, "basi_softmax_ce_fwd_bwd", a.t.data_ptr(), self.label_seg.data_ptr(),
                           C.c_int64(N), c, C.c_double(1.0 / (len(segs) * N)),
                           C.c_float(self.loss_scale / (len(segs) * N)), self.loss_acc.data_ptr(), g.t.data_ptr())
            else:
                self._call(self.lossl, "basi_wbce_sel_fwd_bwd", a.t.data_ptr(), c, 1, self.label_att.data_ptr(),
                           C.c_float(pw), C.c_double(2.0 / (len(segs) * N)),
                           C.c_float(2.0 * self.loss_scale / (len(segs) * N)), C.c_int64(N),
                           self.loss_acc.data_ptr(), g.t.data_ptr())
        self.class_weight = float(cfg.get("class_weight", 0.1))
        for a in clss:
            g = self._grad_of(a)
            a.gw = True
            self._call(self.lossl, "basi_softmax_ce_fwd_bwd", a.t.data_ptr(), self.label_cls.data_ptr(), C.c_int64(B),
                       a.shape[3], C.c_double(1.0 / (len(clss) * B)),
                       C.c_float(self.loss_scale * self.class_weight / (len(clss) * B)),
                       self.loss_acc.data_ptr() + 8, g.t.data_ptr())

    def _lower_loss(self):
        if getattr(self.net, "cascade", False):
            return self._lower_loss_cascade()
        if (self.loss_cfg or {}).get("kind") == "linknet_b" or self._vgg_family:
            return self._lower_loss_linknet()
        cfg = self.loss_cfg
        segname = (cfg or {}).get("seg")
        if segname is None:
            for name in ("conv6_n", "conv6_n_3", "conv6_n_4", "conv6_n_3_coco"):
                if name in self.net.layers:
                    segname = name
        self.seg_name = segname
        self.seg_logits = self._acts[self.net.layers[segname].index]
        clsname = (cfg or {}).get("cls")
        if clsname is None:
            for name in ("class_attention_fc", "class_attention_fc_coco"):
                if name in self.net.layers:
                    clsname = name
        self.cls_name = clsname
        self.cls_logits = self._acts[self.net.layers[clsname].index] if clsname else None
        B, P_h, P_w, nseg = self.seg_logits.shape
        dev = self.device
        self.loss_acc = self._zeros(4, torch.float64)
        self.pred_seg = self._zeros((B, P_h, P_w, 1), torch.int32)
        self.pred_cls = self._zeros((B,), torch.int32) if self.cls_logits is not None else None
        self.post = []
        lp = self.seg_logits.t.data_ptr()
        if nseg == 1:
            self._call(self.post, "basi_threshold", lp, C.c_float(0.5), self.pred_seg.data_ptr(),
                       C.c_int64(B * P_h * P_w))
        else:
            self._call(self.post, "basi_argmax", lp, C.c_int64(B * P_h * P_w), nseg, self.pred_seg.data_ptr())
        if self.cls_logits is not None:
            self._call(self.post, "basi_argmax", self.cls_logits.t.data_ptr(), C.c_int64(B),
                       self.cls_logits.shape[3], self.pred_cls.data_ptr())
        if not self.training:
            return
        assert cfg is not None, "training needs a loss configuration"
        kind = cfg.get("kind", "bce" if nseg == 1 else "softmax")
        N = B * P_h * P_w
        self._pick_loss_scale(N)
        self.lossl = []
        g = self._grad_of(self.seg_logits)
        self.seg_logits.gw = True
        if kind == "bce":
            assert nseg == 1
            self.label_seg = self._zeros((B, P_h, P_w, 1), torch.float32)
            self._call(self.lossl, "basi_wbce_fwd_bwd", lp, self.label_seg.data_ptr(),
                       C.c_float(cfg.get("pos_weight", 3.0)), C.c_double(1.0 / N), C.c_float(self.loss_scale / N),
                       C.c_int64(N), self.loss_acc.data_ptr(), g.t.data_ptr())
        else:
            self.label_seg = self._zeros((B, P_h, P_w, 1), torch.int32)
            self._call(self.lossl, "basi_softmax_ce_fwd_bwd", lp, self.label_seg.data_ptr(), C.c_int64(N), nseg,
                       C.c_double(1.0 / N), C.c_float(self.loss_scale / N), self.loss_acc.data_ptr(), g.t.data_ptr())
        self.class_weight = float(cfg.get("class_weight", 0.2))
        if self.cls_logits is not None:
            self.label_cls = self._zeros((B,), torch.int32)
            gc = self._grad_of(self.cls_logits)
            self.cls_logits.gw = True
            ncls = self.cls_logits.shape[3]
            self._call(self.lossl, "basi_softmax_ce_fwd_bwd", self.cls_logits.t.data_ptr(),
                       self.label_cls.data_ptr(), C.c_int64(B), ncls, C.c_double(1.0 / B),
                       C.c_float(self.loss_scale * self.class_weight / B), self.loss_acc.data_ptr() + 8, gc.t.data_ptr())
        else:
            self.label_cls = None

    # ------------------------------------------------------------------ backward emission
    def _acc_flag(self, act):
        """plan-time: returns 1 if act.grad already holds a contribution, and marks it written."""
        self._grad_of(act)
        f = 1 if act.gw else 0
        act.gw = True
        return f

    def _emit_backward(self):
        # which activations need a gradient: everything except the data input
        self._deferred_bwd = []
        for kind, op in reversed(self._ops):
            self._cur_branch = op.get("branch")
            if self._cur_branch is None and self._deferred_bwd:
                self.bwd.extend(self._deferred_bwd)      # the branches' pooling adjoints, after the branches joined
                self._deferred_bwd = []
            getattr(self, "_bwd_" + kind)(op)
        self._cur_branch = None
        self.bwd.extend(self._deferred_bwd)
        self._deferred_bwd = []
        if self.defer_wgrad:
            self.bwd = [c for c in self.bwd if not c[3].get("side")] + [c for c in self.bwd if c[3].get("side")]

    def _bwd_add(self, op):
        a, b, y = op["a"], op["b"], op["y"]
        if y.grad is None or not y.gw:
            return
        acc_a, acc_b = self._acc_flag(a), self._acc_flag(b)
        self._call(self.bwd, "basi_add_bwd", y.grad.ref, a.grad.ref, acc_a, b.grad.ref, acc_b)

    def _bwd_maxpool2(self, op):
        x, y = op["x"], op["y"]
        if y.grad is None or not y.gw:
            return
        acc = self._acc_flag(x)
        self._call(self.bwd, "basi_maxpool2s2_bwd", y.grad.ref, op["amax"].data_ptr(), x.grad.ref, acc)

    def _bwd_nearest(self, op):
        x, y = op["x"], op["y"]
        if y.grad is None or not y.gw or getattr(x, "no_grad", False):
            return                                   # unused output (finest level) or the click map
        if x.shape[3] == 1 and x.t.dtype == torch.float32:
            self._mask_grad(x)                       # gates: pre-zeroed, every writer adds
            acc = 1
        else:
            acc = self._acc_flag(x)
        self._call(self.bwd, "basi_resize_nearest_bwd", y.grad.ref, x.grad.ref, acc)

    def _bwd_maskmul(self, op):
        feat, mask, y = op["feat"], op["mask"], op["y"]
        if y.grad is None or not y.gw:
            return
        dmask = self._mask_grad(mask)
        acc = self._acc_flag(feat)
        self._call(self.bwd, "basi_mask_mul_bwd", y.grad.ref, feat.ref, mask.t.data_ptr(), 1, 0, feat.grad.ref, acc,
                   dmask.t.data_ptr() if dmask is not None else None)

    def _bwd_softgate(self, op):
        logits, gate = op["logits"], op["gate"]
        if gate.grad is None:
            return
        acc = self._acc_flag(logits)        # (the loss / sigmoid adjoint may or may not have written it already)
        n, h, w, _ = gate.shape
        self._call(self.bwd, "basi_softmax_gate_bwd", logits.t.data_ptr(), gate.grad.t.data_ptr(), C.c_int64(n * h * w),
                   op["C"], op["sel"], C.c_float(op["thr"]), logits.grad.t.data_ptr(), acc)

    def _bwd_relu(self, op):
        x, y = op["x"], op["y"]
        if y.grad is None or not y.gw:
            return
        acc = self._acc_flag(x)
        self._call(self.bwd, "basi_relu_bwd", y.grad.ref, y.ref, x.grad.ref, acc,
                   bytes=self._nbytes(x) * (3 + (1 if acc else 0)))

    def _bwd_sigmoid(self, op):
        x, y = op["x"], op["y"]
        if y.grad is None or not y.gw:
            return
        acc = self._acc_flag(x)
        self._call(self.bwd, "basi_sigmoid_bwd", y.grad.t.data_ptr(), y.t.data_ptr(), x.grad.t.data_ptr(),
                   C.c_int64(op["n"]), acc)

    def _bwd_conv(self, op):
        x, y = op["x"], op["y"]
        dy = y.grad
        if self._vgg_family and (dy is None or not y.gw):
            return                                   # variant B: the finest attention output feeds nothing
        assert dy is not None and y.gw, "conv %s: no gradient reaches its output" % op["name"]
        if op["relu"] and y.dtype == _lib.F32 and y.desc.ld == y.desc.c:
            n = int(np.prod(y.shape))
            self._call(self.bwd, "basi_relu_bwd_f32", dy.t.data_ptr(), y.t.data_ptr(), C.c_int64(n))
        elif op["relu"]:
            # 16-bit (or strided) outputs: mask dy in place; on the tcgen05 path the same pass also reduces dbias
            self._call(self.bwd, "basi_bias_relu_bwd", dy.ref, y.ref, 1,
                       self._gptr(op["b"]) if op.get("tc_bias") else None,
                       writes=[op["b"]] if op.get("tc_bias") else [])
        dptr = C.byref(op["desc"])
        need_dx = x is not self.input
        tc_ok = self.use_tc and x.dtype == _lib.BF16 and y.dtype == _lib.BF16 and (not op["b"] or op.get("tc_bias"))
        lib = _lib.load()
        if op.get("split"):
            # fp32-grade tensor-core path: dy is split once, wgrad and dgrad both read the parts
            if lib.basi_tc_conv_supported_split(_lib.TC_WGRAD, dptr, x.ref, y.ref) == 1:
                self._emit_tc(op, _lib.TC_WGRAD, self.bwd)
            else:
                self._call(self.bwd, "basi_conv_wgrad", dptr, x.ref, dy.ref, self._gptr(op["w"]), None,
                           flops=self._conv_flops(op), writes=[op["w"]], side=True)
            if need_dx:
                acc = self._acc_flag(x)
                if lib.basi_tc_conv_supported_split(_lib.TC_DGRAD, dptr, x.ref, y.ref) == 1:
                    self._emit_tc(op, _lib.TC_DGRAD, self.bwd, acc)
                else:
                    self._call(self.bwd, "basi_conv_dgrad", dptr, dy.ref, self._pptr(op["w"]), x.grad.ref, acc,
                               flops=self._conv_flops(op))
            return
        if tc_ok and lib.basi_tc_conv_supported(_lib.TC_WGRAD, dptr, x.ref, y.ref) == 1:
            self._emit_tc(op, _lib.TC_WGRAD, self.bwd)
        else:
            self._call(self.bwd, "basi_conv_wgrad", dptr, x.ref, dy.ref, self._gptr(op["w"]),
                       self._gptr(op["b"]) if op["b"] else None, flops=self._conv_flops(op),
                       writes=[op["w"]] + ([op["b"]] if op["b"] else []), side=True)
        if need_dx:
            if (tc_ok and self.fuse_bn_dgrad and not x.gw and id(x) in self._bnact_of
                    and lib.basi_tc_conv_supported(_lib.TC_DGRAD, dptr, x.ref, y.ref) == 1
                    and self._try_fused_bn_dgrad(op)):
                return
            acc = self._acc_flag(x)
            if tc_ok and lib.basi_tc_conv_supported(_lib.TC_DGRAD, dptr, x.ref, y.ref) == 1:
                self._emit_tc(op, _lib.TC_DGRAD, self.bwd, acc)
            else:
                self._call(self.bwd, "basi_conv_dgrad", dptr, dy.ref, self._pptr(op["w"]), x.grad.ref, acc,
                           flops=self._conv_flops(op))

    def _try_fused_bn_dgrad(self, op):
        """dgrad of this convolution + the whole BN(+ReLU) backward of the layer that produced its input, one
        cooperative tcgen05 launch (basi_tc_conv_set_bn_bwd): the destination is the gradient wrt that layer's RAW
        conv output, the gradient wrt the activated tensor never exists in memory."""
        lib = _lib.load()
        x = op["x"]
        bn_op, n_cons = self._bnact_of[id(x)]
        if n_cons != 1 or bn_op["res"] is not None or bn_op["res_bn"] is not None:
            return False
        rec = bn_op["main"]
        if rec.x.gw or rec.x.dtype != _lib.BF16 or rec.x.shape != x.shape:
            return False
        dxl = self._grad_of(rec.x)
        handle = C.c_void_p()
        if "w_io" not in op:
            return False
        _lib.call("basi_tc_conv_create", _lib.TC_DGRAD, C.byref(op["desc"]), op["y"].grad.ref, dxl.ref,
                  op["w_io"].data_ptr(), None, 0, C.byref(handle))
        self._tc_plans.append(handle)
        ok = lib.basi_tc_conv_set_bn_bwd(handle, rec.x.ref, rec.bnp.data_ptr(), 1 if bn_op["relu"] else 0, rec.dsums,
                                         C.c_double(rec.count), self._gptr(rec.gamma), self._gptr(rec.beta), rec.cnt_b)
        if ok != 1:
            return False
        rec.x.gw = True
        x.gw = True
        bn_op["bwd_fused"] = True
        self.fused_bn_dgrad += 1
        self.tc_layers += 1
        meta = dict(flops=self._conv_flops(op), layer=op["name"], writes=[rec.gamma, rec.beta],
                    bytes_bn=self._nbytes(rec.x) * 2)
        if getattr(self, "_cur_branch", None) is not None:
            meta["branch"] = self._cur_branch
        self.bwd.append(("basi_tc_conv_run:%d" % _lib.TC_DGRAD, lib.basi_tc_conv_run, (handle,), meta))
        return True

    def _bwd_bnact(self, op):
        if op.get("bwd_fused"):
            return                         # done inside the consumer's dgrad kernel
        out = op["out"]
        dout = out.grad
        assert dout is not None and out.gw
        junction = op["res"] is not None
        # plain BN+ReLU: the mask is recomputed from x (saves reading `out`); junctions need the stored output
        mask = out.ref if (op["relu"] and junction) else None
        from_x = 1 if (op["relu"] and not junction) else 0
        recs = [(op["main"], True)]
        if op["res_bn"] is not None:
            recs.append((op["res_bn"], False))
        for rec, is_main in recs:
            x = rec.x
            dx = self._grad_of(x)
            x.gw = True
            dres, dacc = None, 0
            if is_main and junction and op["res_bn"] is None:
                dacc = self._acc_flag(op["res"])
                dres = op["res"].grad.ref
            nb = self._nbytes(x)
            nin = 3 if mask is not None else 2
            generic = bool(op.get("generic_bwd"))
            if (mask is None and dres is None and self.fuse_bn_bwd and not self.dry_run and not generic
                    and _lib.load().basi_bn_bwd_fused_supported(x.ref) == 1
                    and dout.desc.ld == dout.desc.c and dx.desc.ld == dx.desc.c):
                # reduce + apply in one cooperative launch, dout and x resident in shared memory between the phases
                self._call(self.bwd, "basi_bn_bwd_fused", dout.ref, x.ref, rec.bnp.data_ptr(), from_x, rec.dsums,
                           C.c_double(rec.count), self._gptr(rec.gamma), self._gptr(rec.beta), rec.coef.data_ptr(),
                           rec.cnt_b, dx.ref, bytes=nb * 3, writes=[rec.gamma, rec.beta])
                self.fused_bn_bwd += 1
                continue
            bits_t = op.get("bits")
            if (self.coop_bn_bwd and not self.dry_run and not generic
                    and _lib.load().basi_bn_bwd_coop_supported(x.ref, 1 if (mask is not None and bits_t is None) else 0,
                                                               1 if bits_t is not None else 0, dacc) == 1):
                # reduce + dx pass in one cooperative launch (grid barrier in between)
                nr = 2 + (1.0 / 16 if bits_t is not None else (1 if mask is not None else 0))
                self._call(self.bwd, "basi_bn_bwd_coop", dout.ref, mask if bits_t is None else None,
                           bits_t.data_ptr() if bits_t is not None else None, x.ref, rec.bnp.data_ptr(), from_x,
                           rec.dsums, C.c_double(rec.count), self._gptr(rec.gamma), self._gptr(rec.beta),
                           rec.coef.data_ptr(), rec.cnt_b, dx.ref, dres, dacc,
                           # algorithmic bytes: every input once (the second pass re-reads are not algorithmic)
                           bytes=nb * (nr + 1 + (0 if dres is None else (2 if dacc else 1))),
                           writes=[rec.gamma, rec.beta])
                continue
            if op.get("bits") is not None:
                bits = op["bits"].data_ptr()
                nr = 2 + 1.0 / 16
                self._call(self.bwd, "basi_bn_bwd_reduce_bits", dout.ref, bits, x.ref, rec.bnp.data_ptr(), rec.dsums,
                           C.c_double(rec.count), self._gptr(rec.gamma), self._gptr(rec.beta), rec.coef.data_ptr(),
                           rec.cnt_b, bytes=nb * nr, writes=[rec.gamma, rec.beta])
                self._call(self.bwd, "basi_bn_bwd_apply_bits", dout.ref, bits, x.ref, rec.bnp.data_ptr(),
                           rec.coef.data_ptr(), dx.ref, dres, dacc,
                           bytes=nb * (nr + 1 + (0 if dres is None else (2 if dacc else 1))))
                continue
            self._call(self.bwd, "basi_bn_bwd_reduce", dout.ref, mask, x.ref, rec.bnp.data_ptr(), from_x, rec.dsums,
                       C.c_double(rec.count), self._gptr(rec.gamma), self._gptr(rec.beta), rec.coef.data_ptr(),
                       rec.cnt_b, bytes=nb * nin, writes=[rec.gamma, rec.beta])
            self._call(self.bwd, "basi_bn_bwd_apply", dout.ref, mask, x.ref, rec.bnp.data_ptr(),
                       rec.coef.data_ptr(), from_x, dx.ref, dres, dacc,
                       bytes=nb * (nin + 1 + (0 if dres is None else (2 if dacc else 1))))

    def _bwd_subsample(self, op):
        x, y = op["x"], op["y"]
        if x is self.input:
            return
        acc = self._acc_flag(x)
        self._call(self.bwd, "basi_subsample_bwd", y.grad.ref, op["stride"], x.grad.ref, acc)

    def _bwd_maxpool(self, op):
        x, y = op["x"], op["y"]
        acc = self._acc_flag(x)
        self._call(self.bwd, "basi_maxpool3s2_bwd", y.grad.ref, op["amax"].data_ptr(), x.grad.ref, acc)

    def _bwd_avgpool(self, op):
        x, y = op["x"], op["y"]
        grp = op.get("group")
        if grp is not None:
            # emitted once, when the LAST pool of the group is reached in backward order (all pooled gradients exist
            # by then only after the other branches ran: the call is deferred behind the branch region)
            grp["bwd_seen"] = grp.get("bwd_seen", 0) + 1
            if grp["bwd_seen"] < len(grp["ops"]):
                return
            acc = self._acc_flag(x)
            n, ks, ptrs = self._pool_group_args(grp, True)
            br, self._cur_branch = self._cur_branch, None
            lst = self._deferred_bwd if br is not None else self.bwd
            self._call(lst, "basi_avgpool_multi_bwd", ptrs, n, ks, x.grad.ref, acc, bytes=self._nbytes(x) * 2)
            self._cur_branch = br
            return
        acc = self._acc_flag(x)
        if self._cur_branch is not None:
            # every pyramid branch adds into the SAME gradient tensor (read-modify-write): not on a branch stream, and
            # moved behind the join of the branches
            br, self._cur_branch = self._cur_branch, None
            self._call(self._deferred_bwd, "basi_avgpool_bwd", y.grad.ref, op["k"], x.grad.ref, acc)
            self._cur_branch = br
            return
        self._call(self.bwd, "basi_avgpool_bwd", y.grad.ref, op["k"], x.grad.ref, acc)

    def _bwd_bilinear(self, op):
        x, y = op["x"], op["y"]
        acc = self._acc_flag(x)
        self._call(self.bwd, "basi_bilinear_ac_bwd", y.grad.ref, x.grad.ref, acc)

    def _bwd_gate(self, op):
        feat, logits, y = op["feat"], op["logits"], op["y"]
        acc = self._acc_flag(feat)
        assert logits.gw, "segment-loss gradient must be written before the gate adjoint adds to it"
        self._call(self.bwd, "basi_gate_mul_bwd", y.grad.ref, feat.ref, logits.t.data_ptr(), op["nseg"], op["att"],
                   feat.grad.ref, acc, logits.grad.t.data_ptr())

    def _bwd_skinny(self, op):
        x, y = op["x"], op["y"]
        dy = y.grad
        assert dy is not None and y.gw
        if op["relu"]:
            self._call(self.bwd, "basi_relu_bwd_f32", dy.t.data_ptr(), y.t.data_ptr(), C.c_int64(self.B * op["N"]))
        self._call(self.bwd, "basi_skinny_wgrad", x.t.data_ptr(), x.dtype, C.c_int64(op["lda"]), dy.t.data_ptr(),
                   self._gptr(op["w"]), self._gptr(op["b"]) if op["b"] else None, self.B, op["K"], op["N"],
                   writes=[op["w"]] + ([op["b"]] if op["b"] else []))
        acc = self._acc_flag(x)
        self._call(self.bwd, "basi_skinny_dgrad", dy.t.data_ptr(), self._pptr(op["w"]), x.grad.t.data_ptr(), x.dtype,
                   C.c_int64(op["lda"]), self.B, op["K"], op["N"], acc)

    # ------------------------------------------------------------------ tcgen05 plans
    def _split3(self, act, lst, parts=None):
        """bf16 [hi|mid(|lo)] parts of a float32 activation (parts x C channels); the split kernel is emitted into `lst`
        the first time the tensor is needed (forward inputs stay valid for the weight-gradient pass)."""
        parts = parts or self.split_parts
        key = (id(act), parts)
        if key not in self._split_cache:
            n, h, w, c = act.shape
            a3 = Act(self._zeros((n, h, w, parts * c), torch.bfloat16))
            self._split_cache[key] = (a3, act)            # (keeps `act` alive: the key is its id)
            self._call(lst, "basi_split3_bf16", act.ref, a3.ref, bytes=self._nbytes(act) * 2.5)
        return self._split_cache[key][0]

    def _emit_tc(self, op, kind, lst, acc=0):
        lib = _lib.load()
        off, shape = self.param_index[op["w"]]
        taps, cin, cout = shape[0] * shape[1], shape[2], shape[3]
        split = bool(op.get("split"))
        if "w_io" not in op:
            fp, bp = self.split_parts, self.split_parts_bwd
            kd = lib.basi_tc_split_kcols(cout, bp) if split else cout      # dgrad layout [tap][cin][kd]
            kf = lib.basi_tc_split_kcols(cin, fp) if split else cin        # fprop layout [tap][cout][kf]
            wdt = torch.float16 if (self.precision == "f16" and not split) else torch.bfloat16
            op["w_io"] = self._zeros(taps * cin * kd, wdt)
            op["w_oi"] = self._zeros(taps * cout * kf, wdt)
            self._tc_weights.append((self._pptr(op["w"]), op["w_io"], op["w_oi"], taps, cin, cout,
                                     {(3, 3): 1, (2, 2): 2, (3, 2): 3}[(fp, bp)] if split else 0))
        x, y = op["x"], op["y"]
        handle = C.c_void_p()
        if split:
            if kind == _lib.TC_FPROP:
                x3 = self._split3(x, lst)
                _lib.call("basi_tc_conv_create_split", kind, C.byref(op["desc"]), x3.ref, y.ref, op["w_oi"].data_ptr(),
                          None, 0, self.split_parts, C.byref(handle))
            elif kind == _lib.TC_DGRAD:
                dy3 = self._split3(y.grad, lst, self.split_parts_bwd)
                _lib.call("basi_tc_conv_create_split", kind, C.byref(op["desc"]), dy3.ref, x.grad.ref,
                          op["w_io"].data_ptr(), None, acc, self.split_parts_bwd, C.byref(handle))
            else:
                x3, dy3 = self._split3(x, lst), self._split3(y.grad, lst, self.split_parts_bwd)
                fp, bp = self.split_parts, self.split_parts_bwd
                _lib.call("basi_tc_conv_create_split", kind, C.byref(op["desc"]), x3.ref, dy3.ref, None,
                          self._gptr(op["w"]), 1, fp if fp == bp else 10 * fp + bp, C.byref(handle))
        elif kind == _lib.TC_FPROP:
            _lib.call("basi_tc_conv_create", kind, C.byref(op["desc"]), x.ref, y.ref, op["w_oi"].data_ptr(), None, 0,
                      C.byref(handle))
        elif kind == _lib.TC_DGRAD:
            _lib.call("basi_tc_conv_create", kind, C.byref(op["desc"]), y.grad.ref, x.grad.ref,
                      op["w_io"].data_ptr(), None, acc, C.byref(handle))
        else:
            _lib.call("basi_tc_conv_create", kind, C.byref(op["desc"]), x.ref, y.grad.ref, None,
                      self._gptr(op["w"]), 1, C.byref(handle))
        self._tc_plans.append(handle)
        self.tc_layers += 1
        self._last_tc = handle
        meta = dict(flops=self._conv_flops(op), layer=op["name"])
        if kind == _lib.TC_WGRAD:
            meta["writes"] = [op["w"]]
            meta["side"] = True
        skip = _exp_env("BASI_DEBUG_SKIP_WGRAD")     # timing experiment only (gradients become wrong)
        if kind == _lib.TC_WGRAD and skip and op["name"].startswith(tuple(skip.split(","))):
            return handle
        lst.append(("basi_tc_conv_run:%d" % kind, lib.basi_tc_conv_run, (handle,), meta))
        return handle

    def _refresh_weight_copies(self, stream=None):
        """bf16 copies (both layouts) of every tensor-core layer's weights: one launch for all layers."""
        if not self._tc_weights:
            return
        _lib.use(self.lib_fmt)
        st = stream if stream is not None else torch.cuda.current_stream(self.device).cuda_stream
        if self._pack_table is None:
            entries = (_lib.PackEntry * len(self._tc_weights))()
            blocks = 0
            for i, (wptr, w_io, w_oi, taps, cin, cout, mode) in enumerate(self._tc_weights):
                tco, tci = -(-cout // 32), -(-cin // 32)
                entries[i] = _lib.PackEntry(wptr, w_io.data_ptr(), w_oi.data_ptr(), taps, cin, cout, blocks, tco, tci,
                                            mode, 0)
                blocks += taps * tco * tci
            raw = np.frombuffer(bytes(entries), dtype=np.uint8).copy()
            self._pack_table = torch.from_numpy(raw).to(self.device)
            self._pack_blocks = blocks
        _lib.call("basi_tc_pack_weights_multi", self._pack_table.data_ptr(), len(self._tc_weights), self._pack_blocks,
                  st)

    def __del__(self):
        try:
            lib = _lib.load()
            for h in self._tc_plans:
                lib.basi_tc_conv_destroy(h)
        except Exception:
            pass

    # ------------------------------------------------------------------ execution
    def _run(self, lst, st):
        """Enqueues the calls on stream `st`.  Calls tagged side=True (weight gradients: nothing on the main chain
        depends on them before the optimizer) go to a second stream behind an event, so they overlap the dgrad /
        batch-norm chain; join_side() makes the main stream wait for them."""
        if self.dry_run:
            raise _lib.BasiError("dry_run engine cannot execute")
        side = self._side if self.overlap_wgrad else None
        cur = torch.cuda.current_stream(self.device)
        branches = self._bstreams
        forked, fork_ev = {}, None
        part = self.sm_main > 0 and len(lst) > 0 and lst[0][3].get("bwd")
        if part:        # launch-time grid rules (batch-norm backward, CUDA-core kernels) see the same budget as the plans
            _lib.load().basi_set_sm_budget(self.sm_main, self.sm_wgrad)
        try:
            self._run_calls(lst, st, side, cur, branches, forked, fork_ev)
        finally:
            if part:
                _lib.load().basi_set_sm_budget(0, 0)
        _lib.LAUNCHES += len(lst)

    def _run_calls(self, lst, st, side, cur, branches, forked, fork_ev):
        for name, fn, args, meta in lst:
            b = meta.get("branch") if branches is not None else None
            if b is None and forked:
                self._join_branches(cur, forked)
                forked, fork_ev = {}, None
            if b is not None:
                if not forked:
                    fork_ev = torch.cuda.Event()
                    fork_ev.record(cur)
                if b not in forked:
                    branches[b].wait_event(fork_ev)
                    forked[b] = True
                rc = fn(*args, branches[b].cuda_stream)
            elif side is not None and meta.get("side"):
                side = self._sides[self._side_rr % len(self._sides)]
                self._side_rr += 1
                ev = torch.cuda.Event()
                ev.record(cur)
                side.wait_event(ev)
                rc = fn(*args, side.cuda_stream)
                self._side_dirty = True
            else:
                rc = fn(*args, st)
            if rc != 0:
                raise _lib.BasiError("%s failed (%d): %s" % (name, rc, _lib.last_error()))
        if forked:
            self._join_branches(cur, forked)

    def _join_branches(self, cur, forked):
        for b in forked:
            ev = torch.cuda.Event()
            ev.record(self._bstreams[b])
            cur.wait_event(ev)

    def join_side(self, waiter=None):
        """Makes `waiter` (default: the current stream) wait for the side-stream work issued so far."""
        if self._side is not None and self._side_dirty:
            for sd in self._sides:
                (waiter or torch.cuda.current_stream(self.device)).wait_stream(sd)

    def _stream(self):
        if self.dry_run:
            raise _lib.BasiError("dry_run engine cannot execute")
        return torch.cuda.current_stream(self.device).cuda_stream

    def _zero_step_state(self, st):
        _lib.call("basi_memset", self.bn_scratch.data_ptr(), 0, C.c_int64(self.bn_scratch.numel() * 8), st)
        _lib.call("basi_memset", self.bn_counters.data_ptr(), 0, C.c_int64(self.bn_counters.numel() * 4), st)
        _lib.call("basi_memset", self.loss_acc.data_ptr(), 0, C.c_int64(32), st)
        if self.training:
            _lib.call("basi_memset", self.grads_flat.data_ptr(), 0, C.c_int64(self.n_flat * 4), st)
            for g in self._zero_grads:
                _lib.call("basi_memset", g.t.data_ptr(), 0, C.c_int64(g.t.numel() * g.t.element_size()), st)

    def enable_click_input(self, sigma=30):
        """Device-side batch assembly (A1/A2): uint8 images + clicks -> float32 NHWC4 via basi_clickmap_pack."""
        from .BAISData import click_lut
        B, H, W, _ = self.input.shape
        lut = click_lut((H, W), sigma)
        self.lut_dev = torch.from_numpy(lut).to(self.device)
        self.img_u8 = self._zeros((B, H, W, 3), torch.uint8)
        self.clicks_dev = self._zeros((B, 2), torch.int32)
        self.pre = []
        self._call(self.pre, "basi_clickmap_pack", self.img_u8.data_ptr(), 0, self.clicks_dev.data_ptr(),
                   self.lut_dev.data_ptr(), C.c_int64(lut.size), self.input.t.data_ptr(), B, H, W)

    LABEL_MODES = {"binary": 0, "border": 1, "three": 2, "coco": 3}

    def enable_label_input(self, encoding="binary", target=None, ratio=8):
        """Device-side label decode + click sampling (A3 / F3): uint8 instance-id maps at P x P in, the encoded
        segment labels and the sampled clicks out (basi_label_encode, basi_click_count / _select)."""
        B, P_h, P_w, _ = self.label_seg.shape
        self._lab_mode = self.LABEL_MODES[encoding]
        self._lab_target = (2 if encoding == "coco" else 1) if target is None else int(target)
        self._lab_ratio = int(ratio)
        self.ann_u8 = self._zeros((B, P_h, P_w), torch.uint8)
        self.att_u8 = self._zeros((B, P_h, P_w), torch.uint8) if encoding == "coco" else None
        self.nums_dev = self._zeros((B,), torch.int32)
        self.counts_dev = self._zeros((B,), torch.int32)
        self.k_dev = self._zeros((B,), torch.int32)

    def feed_annotations(self, ann_u8, nums=None, attention_u8=None, rng=None):
        """ann_u8: [B,P,P] uint8 instance-id maps (COCO: uint8 sum of the instance masks), nums: attended instance per
        image (COCO: attention_u8 masks instead).  Encodes the labels on the device, then samples one click per image
        exactly like the reference: k = rng.randint(0, len(np.argwhere(label == target))) drawn on the host in batch
        order (np.random by default), the k-th matching pixel (row-major) times the ratio selected on the device."""
        rng = np.random if rng is None else rng
        _lib.use(self.lib_fmt)
        st = self._stream()
        B, P_h, P_w, _ = self.label_seg.shape
        self.ann_u8.copy_(_as_tensor(ann_u8, torch.uint8).view(self.ann_u8.shape), non_blocking=True)
        if self.att_u8 is not None:
            self.att_u8.copy_(_as_tensor(attention_u8, torch.uint8).view(self.att_u8.shape), non_blocking=True)
        else:
            self.nums_dev.copy_(_as_tensor(nums, torch.int32).view(self.nums_dev.shape), non_blocking=True)
        f32 = self.label_seg.dtype == torch.float32
        _lib.call("basi_label_encode", self.ann_u8.data_ptr(),
                  self.att_u8.data_ptr() if self.att_u8 is not None else None,
                  None if self.att_u8 is not None else self.nums_dev.data_ptr(), self._lab_mode,
                  None if f32 else self.label_seg.data_ptr(), self.label_seg.data_ptr() if f32 else None,
                  B, C.c_int64(P_h * P_w), st)
        _lib.call("basi_click_count", self.label_seg.data_ptr(), 1 if f32 else 0, self._lab_target, B, P_h * P_w,
                  self.counts_dev.data_ptr(), st)
        counts = self.counts_dev.cpu().numpy()
        if np.any(counts <= 0):
            raise ValueError("feed_annotations: an image has no pixel of the attended instance")
        k = np.asarray([rng.randint(0, int(c)) for c in counts], dtype=np.int32)
        self.k_dev.copy_(torch.from_numpy(k), non_blocking=True)
        _lib.call("basi_click_select", self.label_seg.data_ptr(), 1 if f32 else 0, self._lab_target, B, P_h, P_w,
                  self.k_dev.data_ptr(), self._lab_ratio, self.clicks_dev.data_ptr(), st)
        return k

    def feed_clicks(self, images_u8, clicks):
        self.img_u8.copy_(_as_tensor(images_u8, torch.uint8).view(self.img_u8.shape), non_blocking=True)
        self.clicks_dev.copy_(_as_tensor(clicks, torch.int32).view(self.clicks_dev.shape), non_blocking=True)

    def forward_device(self):
        """Forward only, inputs already in self.input (device)."""
        _lib.use(self.lib_fmt)
        st = self._stream()
        self._zero_step_state(st)
        self._run(self.pre, st)
        self._run(self.fwd, st)
        self._run(self.post, st)

    def step_device(self, sync_grads=None):
        """One training step on inputs already resident in the static device buffers."""
        _lib.use(self.lib_fmt)
        st = self._stream()
        self._zero_step_state(st)
        self._run(self.pre, st)
        self._run(self.fwd, st)
        self._run(self.lossl, st)
        self._run(self.post, st)
        if sync_grads is not None and hasattr(sync_grads, "run_backward"):
            sync_grads.run_backward(self, st)          # bucketed all-reduce overlapped with the backward calls
        else:
            self._run(self.bwd, st)
            self.join_side()
            if sync_grads is not None:
                sync_grads(self.grads_flat)
        self.join_side()
        self._side_dirty = False
        if self._train_ranges is None:
            _lib.call("basi_sgd_step", self.params_flat.data_ptr(), self.grads_flat.data_ptr(), self.lr_dev.data_ptr(),
                      C.c_int64(self.n_flat), None, st)
        else:
            # minimize(loss, var_list=...): only the selected variables move (the gradients of the others are
            # computed, like TF does for the shared graph, and ignored)
            for off, n in self._train_ranges:
                _lib.call("basi_sgd_step", self.params_flat.data_ptr() + 4 * off, self.grads_flat.data_ptr() + 4 * off,
                          self.lr_dev.data_ptr(), C.c_int64(n), None, st)
        self._refresh_weight_copies(st)

    def launches_per_step(self):
        n = len(self.pre) + len(self.fwd) + len(self.post)
        if self.training:
            n += len(self.lossl) + len(self.bwd) + 1 + (1 if self._tc_weights else 0)
        return n

    # ---- CUDA graph capture of the whole step
    def capture(self, train=True, sync_grads=None):
        """Captures one whole step (incl. the bucketed NCCL all-reduces when sync_grads is given) into a CUDA graph."""
        torch.cuda.synchronize(self.device)
        lr_saved = self.lr_dev.clone()
        self.lr_dev.zero_()      # the warm-up steps below must not move the weights
        # (a high-priority capture stream for the main chain measured slower: 9.07 vs 8.89 ms)
        s = torch.cuda.Stream(self.device)
        s.wait_stream(torch.cuda.current_stream(self.device))
        with torch.cuda.stream(s):
            for _ in range(2):
                self.step_device(sync_grads) if train else self.forward_device()
        torch.cuda.current_stream(self.device).wait_stream(s)
        torch.cuda.synchronize(self.device)
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=s):
            self.step_device(sync_grads) if train else self.forward_device()
        self.lr_dev.copy_(lr_saved)
        self._graph = g
        return g

    def replay(self):
        self._graph.replay()
        _lib.LAUNCHES += self.launches_per_step()

    # ---- host-facing helpers
    def feed(self, data=None, label_seg=None, label_cls=None, lr=None, mask=None, label_att=None):
        """Copies host arrays (numpy or pinned torch tensors) into the static device buffers."""
        if data is not None:
            self.input.t.copy_(_as_tensor(data, torch.float32).view(self.input.t.shape), non_blocking=True)
        if mask is not None:      # variant B: the click map is a separate 1-channel input
            self.input_mask.t.copy_(_as_tensor(mask, torch.float32).view(self.input_mask.t.shape), non_blocking=True)
        if label_seg is not None:
            self.label_seg.copy_(_as_tensor(label_seg, self.label_seg.dtype).view(self.label_seg.shape),
                                 non_blocking=True)
        if label_att is not None:     # the cascade (F4): {0,1} attention labels next to the 4-class segment labels
            self.label_att.copy_(_as_tensor(label_att, torch.float32).view(self.label_att.shape), non_blocking=True)
        if label_cls is not None and self.label_cls is not None:
            self.label_cls.copy_(_as_tensor(label_cls, torch.int32).view(self.label_cls.shape), non_blocking=True)
        if lr is not None:
            self.lr_dev.fill_(float(lr) / self.loss_scale)      # (gradients carry the loss scale)

    def fetch(self, name, grad=False):
        """Value (or, grad=True, the loss gradient) of a named layer as a float32 NHWC numpy array -- the analogue of
        fetching a symbolic tensor through ``sess.run``.  BN layers fused into a following ReLU / junction return the
        fused output (the tensor the next layer reads)."""
        a = self._acts[self.net.layers[name].index]
        if isinstance(a, tuple):
            if a[0] in ("fused_into_relu", "pre_relu_of"):
                a = a[1]
            else:
                raise KeyError("layer %s is not materialised (%s)" % (name, a[0]))
        if grad:
            if a.grad is None:
                raise KeyError("layer %s has no gradient buffer" % name)
            a = a.grad
        torch.cuda.synchronize(self.device)
        return a.t.float().cpu().numpy()

    # ---- host pipeline of the training loop: prefetch of the next batch, one packed read-back per step
    def stage_batch(self, images_u8, clicks, label_seg=None, label_cls=None, label_att=None):
        """Prefetch: starts the host-to-device copies of the NEXT step's inputs (pinned tensors or numpy arrays) into
        staging buffers on a copy stream, so they overlap the step that is running; commit_staged() moves them into the
        static input buffers of the step plan.  (The reference feeds synchronously through feed_dict,
        back/2AddClass/BAISRunnerTrain.py:158-168; a TF input queue is the closest analogue.)"""
        if not hasattr(self, "_stage"):
            self._copy_stream = torch.cuda.Stream(self.device)
            self._stage = {}
            self._stage_ev = torch.cuda.Event()
            self._commit_ev = None
        srcs = dict(img_u8=(images_u8, getattr(self, "img_u8", None)), clicks=(clicks, getattr(self, "clicks_dev", None)),
                    label_seg=(label_seg, getattr(self, "label_seg", None)),
                    label_cls=(label_cls, getattr(self, "label_cls", None)),
                    label_att=(label_att, getattr(self, "label_att", None)))
        cs = self._copy_stream
        if self._commit_ev is not None:
            cs.wait_event(self._commit_ev)          # the previous commit has read the staging buffers
        self._staged = []
        with torch.cuda.stream(cs):
            for key, (src, dst) in srcs.items():
                if src is None or dst is None:
                    continue
                if key not in self._stage:
                    self._stage[key] = torch.empty_like(dst)
                self._stage[key].copy_(_as_tensor(src, dst.dtype).view(dst.shape), non_blocking=True)
                self._staged.append((self._stage[key], dst))
            self._stage_ev.record(cs)

    def commit_staged(self):
        """Device-to-device copies of the staged batch into the static input buffers, on the current stream."""
        cur = torch.cuda.current_stream(self.device)
        cur.wait_event(self._stage_ev)
        for src, dst in self._staged:
            dst.copy_(src, non_blocking=True)
        self._commit_ev = torch.cuda.Event()
        self._commit_ev.record(cur)
        self._staged = []

    def set_lr(self, lr):
        self.lr_dev.fill_(float(lr) / self.loss_scale)

    def fetch_step(self, extra=()):
        """The per-step read-back of a reference ``sess.run`` fetch list in ONE synchronisation: losses, segment
        logits, predictions (+ class logits / predictions, + `extra` tensors) are copied to pinned host buffers
        asynchronously, the stream is synchronised once, numpy copies are returned."""
        items = [("loss_acc", self.loss_acc), ("raw_output_segment", self.seg_logits.t), ("pred_segment", self.pred_seg)]
        if self.cls_logits is not None:
            items += [("raw_output_classes", self.cls_logits.t), ("pred_classes", self.pred_cls)]
        items += list(extra)
        if not hasattr(self, "_fetch_pinned"):
            self._fetch_pinned = {}
        for name, t in items:
            h = self._fetch_pinned.get(name)
            if h is None or h.shape != t.shape or h.dtype != t.dtype:
                h = self._fetch_pinned[name] = torch.empty(t.shape, dtype=t.dtype, pin_memory=True)
            h.copy_(t, non_blocking=True)
        torch.cuda.current_stream(self.device).synchronize()
        return {name: self._fetch_pinned[name].numpy().copy() for name, _ in items}

    def losses_from(self, acc):
        seg, cls = float(acc[0]), float(acc[1])
        total = seg + (self.class_weight * cls if self.cls_logits is not None else 0.0)
        return total, seg, cls

    def losses(self):
        acc = self.loss_acc.cpu().numpy()
        seg, cls = float(acc[0]), float(acc[1])
        total = seg + (self.class_weight * cls if self.cls_logits is not None else 0.0)
        return total, seg, cls


def _as_tensor(x, dtype):
    if isinstance(x, torch.Tensor):
        return x if x.dtype == dtype else x.to(dtype)
    a = np.asarray(x)
    return torch.from_numpy(np.ascontiguousarray(a)).to(dtype)
