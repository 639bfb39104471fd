"""CPU oracle for the BAIS PSPNet hot path -- TEST INFRASTRUCTURE ONLY.

This module is a from-the-source restatement (numpy + PyTorch-CPU) of what the
reference computes on its hot path.  It is the *checker*: only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl
reference`` legs may import it.  Nothing under ``instance-segment-basi_b200/``
(the product) imports it.

PARITY -- what is pinned against the reference itself, and how:

* DATA functions (``mask_gaussian``, ``pack_input``, ``encode_labels_binary`` / ``_three`` / ``_border``,
  ``sample_click``): bit for bit against outputs of the reference's own BAISData.py code, which is numpy + PIL and runs
  in the build container (tests/golden/make_reference_golden.py -> reference_data.npz -> tests/test_reference_golden.py).
* CONVOLUTION PADDING / STRIDE SEMANTICS (``conv2d`` 'SAME' at stride 1 and 2 on even and odd inputs, explicit padding +
  'VALID', strided 1x1 = subsample, ``tf_same_pad``): against the numeric golden vectors of the reference's own
  vendored TF-slim tests (slim/nets/resnet_v1_test.py:58-153 -> tests/golden/slim_reference_tests.json).
* NETWORK STRUCTURE, LOSS COMPOSITION, LEARNING-RATE FORMULA, OPTIMIZER VARIABLE LISTS of every network here
  (``pspnet_forward`` / ``losses`` / ``train_step`` for 1NoClass, 2AddClass, 3ThreeClass, 4BorderClass, 5COCO;
  ``attention_u_*``; ``linknet_b_*``; ``linknet_top_*``): against the reference's OWN Python executed in the build
  container.  tests/golden/make_reference_net_golden.py imports the unmodified reference modules over an eager float64
  stand-in for the ~60 ``tf.*`` / ``slim.*`` calls they make (tests/golden/tf1_shim) and runs the reference's own
  ``Train.build_net()`` / whole ``Train.__init__`` (for 8AttentionU, HEAD and 90AttentionSingle2 including the reference's
  own Data reader on tests/golden/voc_mini and slim's vgg_16 through nets_factory); tests/test_reference_net_golden.py
  holds this oracle to what that produced -- every exposed layer, logits, predictions, losses, learning rate, every
  gradient and the SGD update, to float64 round-off (1e-9 ... 1e-15).  Two discrepancies were FOUND that way and fixed:
  the hand-unrolled trunk of back/8AttentionU/BAISNet.py feeds the PRE-ReLU junction sum to 31 convolutions
  (``_pspnet_trunk(wiring="8AttentionU")``), and variant B's base learning rate is 5e-4.
* What that does NOT pin: the arithmetic of each TensorFlow primitive.  TensorFlow 1.x is third-party, un-vendored and
  un-pinned (no requirements file; TF1 is implied by ``tf.contrib.slim`` / ``tf.placeholder``) and cannot be installed
  here; the stand-in's primitives are a second, independently written restatement of the published TF1 op semantics
  (NHWC tap-sum convolutions, window stacks, explicit bilinear matrices, float32 index arithmetic of
  resize_nearest_neighbor) that agrees with this module's to round-off.  For batch norm, pooling, resize, gating and the
  loss formulas "PARITY UNPINNED against TensorFlow's kernels" therefore still holds: they follow the reference's call
  sites plus the published op definitions and are additionally held by hand-computed known-answer cases and fp64
  finite-difference checks (tests/test_oracle.py).

What each function restates (paths relative to /root/reference):

* ``mask_gaussian``          back/2AddClass/BAISData.py:189-202 (sigma=30),
                             back/5COCO/BAISData.py:405-417 (sigma=20)
* ``pack_input``             back/2AddClass/BAISData.py:79-80
* ``encode_labels_border``   back/4BorderClass/BAISData.py:143-165 (``_three``: the has_255=False branch,
                             ``_coco``: back/5COCO/BAISData.py:361-369, ``sample_click``: 2AddClass :63-66)
* ``conv2d`` / ``atrous``    back/2AddClass/BAISPSPNet.py:118-146 (tf.nn.conv2d /
                             tf.pad + tf.nn.atrous_conv2d, NHWC, HWIO weights)
* ``batch_norm``             back/2AddClass/BAISPSPNet.py:204-236 (training=True always)
* ``max_pool_3x3_s2_same``   back/2AddClass/BAISPSPNet.py:152-155, used at :269
* ``avg_pool``               back/2AddClass/BAISPSPNet.py:157-160
* ``resize_bilinear_ac``     back/2AddClass/BAISPSPNet.py:242-244 (align_corners=True)
* ``resize_bilinear_legacy`` back/4BorderClass/BAISRunnerGUI.py:30 (TF1 default)
* ``pspnet_forward``         back/2AddClass/BAISPSPNet.py:260-734 and the head
                             variants 4BorderClass/BAISPSPNet.py:716-737,
                             5COCO/BAISPSPNet.py:716-737
* ``losses`` / ``train_step``back/{1NoClass,2AddClass,4BorderClass,5COCO}/BAISRunnerTrain.py
                             build_net (loss, poly LR, plain SGD)
* ``linknet_b_*``            variant B: slim/nets/vgg.py:187-196 (vgg_16 trunk), back/90AttentionSingle2/BAISNet.py
                             :653-810 (LinkNet._attention / _classifies / build), BAISRunnerTrain.py:116-156 (cal_loss)
* ``linknet_top_*``          the current-HEAD path: BAISNet.py:117-269 (LinkNet._feature / _decoder / _segment / build),
                             BAISRunnerTrain.py:97-115 (cal_loss, pos_weight 1)
* ``pspnet_forward_rounded`` the same forward with 16-bit storage roundings injected (no reference counterpart:
                             the noise model the bf16 CUDA path is measured against)
* ``predict_*``              back/2AddClass/BAISRunnerTrain.py:88-89,
                             back/4BorderClass/BAISRunnerOne.py:40-45,
                             back/4BorderClass/BAISRunnerGUI.py:29-34,84
"""
from __future__ import annotations

import math
from collections import OrderedDict

import numpy as np
import torch
import torch.nn.functional as F

BN_EPS = 1e-5
# when set to a list, every ReLU input appends min |pre-activation| (used to keep golden fixtures away from
# float32 ReLU ties: an input within one float32 ulp of zero makes any float32 run flip that mask element)
TRACE_RELU_MARGIN = None


def _relu(t):
    if TRACE_RELU_MARGIN is not None:
        TRACE_RELU_MARGIN.append(float(t.detach().abs().min()))
    return F.relu(t)

# --------------------------------------------------------------------------
# A1 / A2 / A3: host-side data path
# --------------------------------------------------------------------------


def mask_gaussian(image_size, where, sigma=30):
    """Gaussian click map, float64 math rounded once to float32."""
    x = np.arange(0, image_size[1], 1, float)
    y = np.arange(0, image_size[0], 1, float)[:, np.newaxis]
    x0, y0 = where[1], where[0]
    return np.exp(-4 * np.log(2) * ((x - x0) ** 2 + (y - y0) ** 2) / sigma ** 2).astype(np.float32)


def pack_input(image_u8, where, sigma=30):
    """uint8 HxWx3 image + click -> float32 HxWx4 (image/255 in float32, click map)."""
    img = np.asarray(image_u8, dtype=np.float32)
    img /= 255
    m = mask_gaussian(img.shape[:2], where, sigma)
    return np.concatenate((img, np.expand_dims(m, 2)), 2)


def encode_labels_border(ann_u8, num):
    """4-class label map of the 4BorderClass snapshot (has_255=True branch).

    ann_u8: uint8 instance-id map (0 background, 255 border, k instance k).
    Result: 0 other instance, 1 the attended instance, 2 border, 3 background --
    including the uint8 wrap-around of (0 - 1) // 84 == 3.
    """
    ann = np.asarray(ann_u8, dtype=np.uint8)
    ann = np.where(ann == 255, 171, ann).astype(np.uint8)
    one = np.where(ann == num, 85, ann).astype(np.uint8)
    return ((one - np.uint8(1)) // np.uint8(84)).astype(np.uint8)


def encode_labels_binary(ann_u8, num):
    return np.where(np.asarray(ann_u8) == num, 1, 0)


def encode_labels_three(ann_u8, num):
    """3-class label map (has_255=False branch of back/4BorderClass/BAISData.py:146-160; the 3ThreeClass snapshot):
    border -> background, 0 other instance, 1 attended instance, 2 background (uint8 wrap of (0 - 1) // 127)."""
    ann = np.asarray(ann_u8, dtype=np.uint8)
    ann = np.where(ann == 255, 0, ann).astype(np.uint8)
    one = np.where(ann == num, 128, ann).astype(np.uint8)
    return ((one - np.uint8(1)) // np.uint8(127)).astype(np.uint8)


def encode_labels_coco(ann_sum_u8, attention_u8):
    """COCO label map (back/5COCO/BAISData.py:361-369): 0 background, 1 other instance, 2 attended instance."""
    lab = np.where(np.asarray(ann_sum_u8) > 0, 1, 0)
    return np.where(np.asarray(attention_u8) > 0, 2, lab)


def sample_click(label_map, k, ratio=8, target=1):
    """where = np.argwhere(ann == target)[k] * ratio (back/2AddClass/BAISData.py:63-66 with the drawn index k)."""
    where = np.argwhere(np.asarray(label_map) == target)
    return [int(where[k][0]) * ratio, int(where[k][1]) * ratio]


# --------------------------------------------------------------------------
# TF1 op semantics on NCHW torch tensors
# --------------------------------------------------------------------------


def tf_same_pad(n_in, k, s, d=1):
    keff = (k - 1) * d + 1
    n_out = -(-n_in // s)
    total = max((n_out - 1) * s + keff - n_in, 0)
    return total // 2, total - total // 2


def conv2d(x, w_hwio, stride=1, padding="VALID", dilation=1, bias=None):
    """tf.nn.conv2d / atrous_conv2d.  x: NCHW, w: [kh,kw,Cin,Cout]."""
    w = w_hwio.permute(3, 2, 0, 1)
    kh, kw = w_hwio.shape[0], w_hwio.shape[1]
    if padding == "SAME":
        pt, pb = tf_same_pad(x.shape[2], kh, stride, dilation)
        pl, pr = tf_same_pad(x.shape[3], kw, stride, dilation)
        x = F.pad(x, (pl, pr, pt, pb))
    elif isinstance(padding, int):
        x = F.pad(x, (padding,) * 4)
    return F.conv2d(x, w, bias, stride=stride, dilation=dilation)


def batch_norm(x, gamma, beta, relu=False, eps=BN_EPS):
    """tf.layers.batch_normalization(training=True): batch mean, biased variance."""
    mean = x.mean(dim=(0, 2, 3), keepdim=True)
    var = ((x - mean) ** 2).mean(dim=(0, 2, 3), keepdim=True)
    y = (x - mean) * torch.rsqrt(var + eps) * gamma.view(1, -1, 1, 1) + beta.view(1, -1, 1, 1)
    return _relu(y) if relu else y


def max_pool_3x3_s2_same(x):
    pt, pb = tf_same_pad(x.shape[2], 3, 2)
    pl, pr = tf_same_pad(x.shape[3], 3, 2)
    x = F.pad(x, (pl, pr, pt, pb), value=float("-inf"))
    return F.max_pool2d(x, 3, 2)


def avg_pool(x, k):
    return F.avg_pool2d(x, k, k)


def resize_bilinear_ac(x, size):
    return F.interpolate(x, size=tuple(size), mode="bilinear", align_corners=True)


def resize_bilinear_legacy(x, size):
    """TF1 resize_bilinear(align_corners=False): src = dst * in/out, no half-pixel offset."""
    n, c, h, w = x.shape
    oh, ow = size

    def axis(n_in, n_out):
        src = torch.arange(n_out, dtype=x.dtype) * (n_in / n_out)
        lo = src.floor().long().clamp(max=n_in - 1)
        hi = (lo + 1).clamp(max=n_in - 1)
        return lo, hi, src - lo.to(x.dtype)

    y0, y1, fy = axis(h, oh)
    x0, x1, fx = axis(w, ow)
    top = x[:, :, y0][:, :, :, x0] * (1 - fx) + x[:, :, y0][:, :, :, x1] * fx
    bot = x[:, :, y1][:, :, :, x0] * (1 - fx) + x[:, :, y1][:, :, :, x1] * fx
    return top * (1 - fy).view(1, 1, -1, 1) + bot * fy.view(1, 1, -1, 1)


def resize_nearest(x, size):
    n, c, h, w = x.shape
    oh, ow = size
    yi = torch.clamp((torch.arange(oh) * (h / oh)).floor().long(), max=h - 1)
    xi = torch.clamp((torch.arange(ow) * (w / ow)).floor().long(), max=w - 1)
    return x[:, :, yi][:, :, :, xi]


def mask_multiply(feature_nchw, mask_n1hw):
    """tf.multiply(feature, mask): the click / attention map broadcast over channels
    (back/7OLD/BAISNet.py:535, back/90AttentionSingle2/BAISNet.py:748,773-776)."""
    return feature_nchw * mask_n1hw


def attention_gate(logits_nchw, sel=1, thr=0.9):
    """softmax over channels -> channel `sel` -> tf.where(p > thr, p, 0)
    (back/90AttentionSingle2/BAISNet.py:743-746; thr < 0: plain softmax channel, back/8AttentionU/BAISNet.py:586)."""
    p = torch.softmax(logits_nchw, dim=1)[:, sel:sel + 1]
    return torch.where(p > thr, p, torch.zeros_like(p))


def weighted_cross_entropy_with_logits(targets, logits, pos_weight):
    z, x, q = targets, logits, pos_weight
    return (1 - z) * x + (1 + (q - 1) * z) * (torch.log1p(torch.exp(-x.abs())) + F.relu(-x))


# --------------------------------------------------------------------------
# Variants, parameter inventory, init
# --------------------------------------------------------------------------

VARIANTS = {
    # name: (segment head layer, class fc layer, has class head, seg loss kind)
    "1NoClass": ("conv6_n", None, False, "bce"),
    "2AddClass": ("conv6_n", "class_attention_fc", True, "bce"),
    "3ThreeClass": ("conv6_n_3", "class_attention_fc", True, "softmax"),
    "4BorderClass": ("conv6_n_4", "class_attention_fc", True, "softmax"),
    "5COCO": ("conv6_n_3_coco", "class_attention_fc_coco", True, "softmax"),
}

STAGES = ((2, 3, 1, 1, 1), (3, 4, 2, 2, 1), (4, 23, 4, 1, 2), (5, 3, 8, 1, 4))  # (stage, blocks, mid/F, stride, dilation)
PSP_LEVELS = (1, 2, 3, 6)


def param_specs(variant="2AddClass", num_classes=21, num_segment=1, filter_number=32):
    """Ordered {tf variable name: shape} for the PSPNet lineage (HWIO conv weights)."""
    seg_name, fc_name, has_class, _ = VARIANTS[variant]
    Fn = filter_number
    specs = OrderedDict()

    def conv(name, k, cin, cout, biased=False):
        specs[name + "/weights"] = (k, k, cin, cout)
        if biased:
            specs[name + "/biases"] = (cout,)

    def bn(name, c):
        specs["%s/%s/gamma" % (name, name)] = (c,)
        specs["%s/%s/beta" % (name, name)] = (c,)

    conv("conv1_1_3x3_s2_n", 3, 4, Fn); bn("conv1_1_3x3_s2_bn", Fn)
    conv("conv1_2_3x3", 3, Fn, Fn); bn("conv1_2_3x3_bn", Fn)
    conv("conv1_3_3x3", 3, Fn, 2 * Fn); bn("conv1_3_3x3_bn", 2 * Fn)
    cin = 2 * Fn
    for stage, blocks, mult, _, _ in STAGES:
        mid, out = mult * Fn, 4 * mult * Fn
        for b in range(1, blocks + 1):
            p = "conv%d_%d" % (stage, b)
            if b == 1:
                conv(p + "_1x1_proj", 1, cin, out); bn(p + "_1x1_proj_bn", out)
            conv(p + "_1x1_reduce", 1, cin, mid); bn(p + "_1x1_reduce_bn", mid)
            conv(p + "_3x3", 3, mid, mid); bn(p + "_3x3_bn", mid)
            conv(p + "_1x1_increase", 1, mid, out); bn(p + "_1x1_increase_bn", out)
            cin = out
    psp = cin // 4
    for lvl in PSP_LEVELS:
        conv("conv5_3_pool%d_conv" % lvl, 1, cin, psp); bn("conv5_3_pool%d_conv_bn" % lvl, psp)
    conv("conv5_4", 3, 2 * cin, psp); bn("conv5_4_bn", psp)
    conv(seg_name, 1, psp, num_segment, biased=True)
    if has_class:
        conv("class_attention_conv", 5, cin, 16 * Fn, biased=True)
        specs[fc_name + "/weights"] = (16 * Fn, num_classes)
        specs[fc_name + "/biases"] = (num_classes,)
    return specs


def init_params(specs, seed=0, dtype=np.float32, trained_like=False):
    """glorot_uniform for weights and biases (tf.get_variable default), gamma=1, beta=0.

    trained_like=True perturbs gamma/beta so that parity tests do not run on the
    degenerate gamma=1/beta=0 point (SURVEY section 7 'hard parts'), and keeps the residual
    branches small (increase_bn gamma in [0.05, 0.25]) the way trained residual nets are: a
    random-init 100-layer batch-stat-BN net is chaotic (float32 vs float64 runs of this very
    oracle differ by 1e-3 on the logits), which would make any parity tolerance meaningless.
    """
    rng = np.random.RandomState(seed)
    out = OrderedDict()
    for name, shape in specs.items():
        if name.endswith("/gamma"):
            if not trained_like:
                v = np.ones(shape)
            elif name.endswith("increase_bn/gamma"):
                v = rng.uniform(0.05, 0.25, shape)
            else:
                v = rng.uniform(0.5, 1.5, shape)
        elif name.endswith("/beta"):
            v = np.zeros(shape) if not trained_like else rng.uniform(-0.3, 0.3, shape)
        else:
            if len(shape) == 4:
                fan_in, fan_out = shape[0] * shape[1] * shape[2], shape[0] * shape[1] * shape[3]
            elif len(shape) == 2:
                fan_in, fan_out = shape
            else:
                fan_in = fan_out = shape[0]
            lim = math.sqrt(6.0 / (fan_in + fan_out))
            v = rng.uniform(-lim, lim, shape)
        out[name] = np.ascontiguousarray(v, dtype=dtype)
    return out


# --------------------------------------------------------------------------
# A4..A11: network forward
# --------------------------------------------------------------------------


def _pspnet_trunk(params, x, L, wiring="2AddClass"):
    """conv1_1 .. conv5_3/relu of the half-width dilated ResNet-101 (NCHW in, NCHW out); fills L with the named layers.

    wiring="8AttentionU": back/8AttentionU/BAISNet.py:133-480 unrolls the same trunk by hand, and there the variable
    `net_input` still holds the junction SUM when the next convolution is built (:163-165 ``net_input = Net.add(...)``,
    ``net_input_conv2_2_relu = Net.relu(net_input, ...)``, ``Net.conv(net_input, ... 'conv2_3_1x1_reduce')``): the
    1x1_reduce of every non-entry block except conv2_2 and the 1x1_proj of conv3_1 / conv4_1 / conv5_1 read the
    PRE-ReLU sum of the previous junction; the shortcuts, conv2_2_1x1_reduce and the 1x1_reduce of the entry blocks read
    its ReLU.  Pinned by running that file (tests/golden/reference_net_8AttentionU.json, key "wiring")."""

    def W(n):
        return params[n + "/weights"]

    def BN(x, n, relu):
        return batch_norm(x, params["%s/%s/gamma" % (n, n)], params["%s/%s/beta" % (n, n)], relu)

    x = _relu(BN(conv2d(x, W("conv1_1_3x3_s2_n"), 2, "SAME"), "conv1_1_3x3_s2_bn", False))
    x = BN(conv2d(x, W("conv1_2_3x3"), 1, "SAME"), "conv1_2_3x3_bn", True)
    x = BN(conv2d(x, W("conv1_3_3x3"), 1, "SAME"), "conv1_3_3x3_bn", True)
    L["conv1_3_3x3_bn"] = x
    x = max_pool_3x3_s2_same(x)
    L["pool1_3x3_s2"] = x
    pre_sum = None                                          # the previous junction before its ReLU
    for stage, blocks, _, stride, dil in STAGES:
        for b in range(1, blocks + 1):
            p = "conv%d_%d" % (stage, b)
            s = stride if b == 1 else 1
            reads_sum = wiring == "8AttentionU" and pre_sum is not None and p != "conv2_2"
            red_in = x
            if b == 1:
                sc = BN(conv2d(pre_sum if reads_sum else x, W(p + "_1x1_proj"), s), p + "_1x1_proj_bn", False)
            else:
                sc = x
                if reads_sum:
                    red_in = pre_sum
            def step(t, conv_name, relu, *conv_args):
                c = conv2d(t, W(conv_name), *conv_args)
                L[conv_name] = c
                o = BN(c, conv_name + "_bn", relu)
                L[conv_name + "_bn"] = o
                return o

            y = step(red_in, p + "_1x1_reduce", True, s)
            y = step(y, p + "_3x3", True, 1, dil, dil)                            # tf.pad(d) + (atrous) VALID
            y = step(y, p + "_1x1_increase", False, 1)
            pre = sc + y
            pre_sum = pre
            x = _relu(pre)
            L[p] = pre
            L[p + "/relu"] = x
    return x


def pspnet_forward(params, data_nhwc, variant="2AddClass", num_segment=1, last_pool_size=40,
                   attention_channel=None, keep=()):
    """Forward pass.  params: {tf name: torch tensor}; data: [B,S,S,4] torch tensor (NHWC).

    Returns dict with NHWC tensors: the segment logits under the variant's head
    name, class logits (if any) and any layer named in ``keep``.
    """
    seg_name, fc_name, has_class, _ = VARIANTS[variant]
    if attention_channel is None:
        attention_channel = {"3ThreeClass": 1, "4BorderClass": 1, "5COCO": 2}.get(variant, 0)
    L = {}

    def W(n):
        return params[n + "/weights"]

    def BN(x, n, relu):
        return batch_norm(x, params["%s/%s/gamma" % (n, n)], params["%s/%s/beta" % (n, n)], relu)

    c53 = _pspnet_trunk(params, data_nhwc.permute(0, 3, 1, 2), L)
    P = last_pool_size
    size = c53.shape[2:4]
    branches = {}
    for lvl in PSP_LEVELS:
        k = P // lvl
        n = "conv5_3_pool%d" % lvl
        y = avg_pool(c53, k)
        y = BN(conv2d(y, W(n + "_conv"), 1), n + "_conv_bn", True)
        branches[lvl] = resize_bilinear_ac(y, size)
        L[n + "_interp"] = branches[lvl]
    cat = torch.cat([c53, branches[6], branches[3], branches[2], branches[1]], dim=1)
    y = BN(conv2d(cat, W("conv5_4"), 1, "SAME"), "conv5_4_bn", True)
    L["conv5_4_bn"] = y
    seg = conv2d(y, W(seg_name), 1, bias=params[seg_name + "/biases"])
    L[seg_name] = seg
    if has_class:
        gate = seg if num_segment == 1 else seg[:, attention_channel:attention_channel + 1]
        m = F.relu(L["conv5_3"]) * gate
        L["class_attention_multiply"] = m
        m = avg_pool(m, P // 5)
        m = F.relu(conv2d(m, W("class_attention_conv"), 5, bias=params["class_attention_conv/biases"]))
        assert m.shape[2] == 1 and m.shape[3] == 1, "class head expects a 1x1 map after the 5x5/s5 conv"
        m = m[:, :, 0, 0]
        L[fc_name] = m @ params[fc_name + "/weights"] + params[fc_name + "/biases"]
    out = {seg_name: L[seg_name].permute(0, 2, 3, 1)}
    if has_class:
        out[fc_name] = L[fc_name]
    for k_ in keep:
        v = L[k_]
        out[k_] = v.permute(0, 2, 3, 1) if v.dim() == 4 else v
    out["__nchw__"] = {k_: L[k_] for k_ in keep}     # graph tensors (for activation gradients)
    return out


# --------------------------------------------------------------------------
# F1: variant B -- slim vgg_16 trunk + click-gated attention cascade (back/90AttentionSingle2/BAISNet.py:653-810)
# --------------------------------------------------------------------------

VGG_BLOCKS = ((1, 2, 64), (2, 2, 128), (3, 3, 256), (4, 3, 512), (5, 3, 512))      # slim/nets/vgg.py:187-196


def linknet_b_specs(num_classes=21, width=1.0):
    """Ordered {tf variable name: shape}.  width scales the vgg_16 channel counts (1.0 = the reference; tests use
    narrower trunks).  Attention / decoder scopes as in LinkNet.build (:756-810) and _attention / _classifies."""
    specs = OrderedDict()

    def conv(name, k, cin, cout, biased):
        specs[name + "/weights"] = (k, k, cin, cout)
        if biased:
            specs[name + "/biases"] = (cout,)

    def bn(name, c):      # tf.layers.batch_normalization(name=leaf) inside variable_scope(leaf): doubled leaf scope
        leaf = name.split("/")[-1]
        specs["%s/%s/gamma" % (name, leaf)] = (c,)
        specs["%s/%s/beta" % (name, leaf)] = (c,)

    cin = 3
    ch = {}
    for blk, reps, c in VGG_BLOCKS:
        c = max(8, int(c * width))
        for r in range(1, reps + 1):
            conv("vgg_16/conv%d/conv%d_%d" % (blk, blk, r), 3, cin, c, True)
            cin = c
        ch[blk] = c
    c1, c2, c3, c4 = ch[2], ch[3], ch[4], ch[5]          # block1..block4 = conv2_2, conv3_3, conv4_3, conv5_3
    for lvl, cin_a, c_in_size, c_out in ((4, 2 * c4, c4, c3), (3, c3 + c3, c3, c2), (2, c2 + c2, c2, c1),
                                         (1, c1 + c1, c1, c1)):
        sc = "attention_%d/attention_%d_attention" % (lvl, lvl)
        conv(sc + "/a_conv_1", 3, cin_a, c_in_size // 4, False); bn(sc + "/a_conv_1_bn", c_in_size // 4)
        conv(sc + "/a_conv_2", 3, c_in_size // 4, 32, False); bn(sc + "/a_conv_2_bn", 32)
        conv(sc + "/a_conv_3", 1, 32, 2, True)
        conv(sc + "/a_conv_o", 1, cin_a, c_out, False)
        if lvl == 4:
            dc = "attention_4/segment_attention_4_decoder"
            conv(dc + "/d_c_conv_1", 3, c3, c4, False); bn(dc + "/d_c_conv_1_bn", c4)
            conv(dc + "/d_c_conv_2", 3, c4, 2 * c4, False); bn(dc + "/d_c_conv_2_bn", 2 * c4)
            conv(dc + "/class_attention_conv", 3, 2 * c4, 4 * c4, True)
            specs[dc + "/class_attention_fc/weights"] = (4 * c4, num_classes)
            specs[dc + "/class_attention_fc/biases"] = (num_classes,)
    return specs


def linknet_b_forward(params, image_nhwc, mask_nhw1, p_size=None, thr=0.9):
    """LinkNet.build of snapshot 90AttentionSingle2 (:756-810): returns (attention logits [4 x NHWC, coarse -> fine],
    class logits).  p_size: class-head pooling window (reference 15 at S = 720; default block4 size // 3)."""
    def W(n):
        return params[n + "/weights"]

    def BN(x, n, relu):
        leaf = n.split("/")[-1]
        return batch_norm(x, params["%s/%s/gamma" % (n, leaf)], params["%s/%s/beta" % (n, leaf)], relu)

    x = image_nhwc.permute(0, 3, 1, 2)
    mask = mask_nhw1.permute(0, 3, 1, 2)
    blocks = {}
    for blk, reps, _ in VGG_BLOCKS:
        for r in range(1, reps + 1):
            n = "vgg_16/conv%d/conv%d_%d" % (blk, blk, r)
            x = F.relu(conv2d(x, W(n), 1, "SAME", bias=params[n + "/biases"]))
        blocks[blk] = x
        if blk < 5:
            x = F.max_pool2d(x, 2, 2)
    b1, b2, b3, b4 = blocks[2], blocks[3], blocks[4], blocks[5]
    attentions = []
    p = resize_nearest(mask, b4.shape[2:4])
    f4 = mask_multiply(b4, p)
    feat = torch.cat([f4, f4], dim=1)
    cls = None
    up = None
    for lvl, blk, nxt in ((4, b4, b3), (3, b3, b2), (2, b2, b1), (1, b1, b1)):
        if lvl != 4:
            feat = torch.cat([mask_multiply(blk, p), up], dim=1)
        sc = "attention_%d/attention_%d_attention" % (lvl, lvl)
        a = BN(conv2d(feat, W(sc + "/a_conv_1"), 1, "SAME"), sc + "/a_conv_1_bn", True)
        a = BN(conv2d(a, W(sc + "/a_conv_2"), 1, "SAME"), sc + "/a_conv_2_bn", True)
        logit = conv2d(a, W(sc + "/a_conv_3"), 1, bias=params[sc + "/a_conv_3/biases"])
        attentions.append(logit.permute(0, 2, 3, 1))
        gate = attention_gate(logit, 1, thr)
        out = conv2d(mask_multiply(feat, gate), W(sc + "/a_conv_o"), 1)
        if lvl == 4:
            dc = "attention_4/segment_attention_4_decoder"
            c = BN(conv2d(out, W(dc + "/d_c_conv_1"), 1, "SAME"), dc + "/d_c_conv_1_bn", True)
            c = BN(conv2d(c, W(dc + "/d_c_conv_2"), 1, "SAME"), dc + "/d_c_conv_2_bn", True)
            ps = p_size if p_size is not None else c.shape[2] // 3
            c = avg_pool(c, ps)
            c = F.relu(conv2d(c, W(dc + "/class_attention_conv"), 3, bias=params[dc + "/class_attention_conv/biases"]))
            assert c.shape[2] == 1 and c.shape[3] == 1
            cls = c[:, :, 0, 0] @ params[dc + "/class_attention_fc/weights"] + params[dc + "/class_attention_fc/biases"]
        up = resize_nearest(out, nxt.shape[2:4])
        p = resize_nearest(gate, nxt.shape[2:4])
    return attentions, cls


def linknet_b_losses(attentions, cls_logits, label_seg_nhw1, label_cls, pos_weight=3.0):
    """cal_loss of back/90AttentionSingle2/BAISRunnerTrain.py:116-156: per attention map the labels are nearest-resized,
    one-hot (depth 2) and scored with weighted_cross_entropy_with_logits(pos_weight=3) averaged over ALL 2N elements;
    loss = mean over the maps + mean softmax CE of the class head."""
    lab = label_seg_nhw1.permute(0, 3, 1, 2).to(attentions[0].dtype)
    terms = []
    for a in attentions:
        z = resize_nearest(lab, a.shape[1:3]).permute(0, 2, 3, 1).reshape(-1).long()
        t = F.one_hot(z, 2).to(a.dtype)
        terms.append(weighted_cross_entropy_with_logits(t, a.reshape(-1, 2), pos_weight).mean())
    loss_att = sum(terms) / len(terms)
    loss_cls = F.cross_entropy(cls_logits, label_cls.long(), reduction="mean")
    return loss_att + loss_cls, loss_att, loss_cls


def linknet_b_train_step(params_np, image, mask, label_seg, label_cls, lr=5e-3, dtype=torch.float64, p_size=None,
                         pos_weight=3.0):
    p = to_torch(params_np, dtype, requires_grad=True)
    att, cls = linknet_b_forward(p, torch.as_tensor(np.asarray(image)).to(dtype),
                                 torch.as_tensor(np.asarray(mask)).to(dtype), p_size)
    loss, la, lc = linknet_b_losses(att, cls, torch.as_tensor(np.asarray(label_seg)), torch.as_tensor(np.asarray(label_cls)),
                                    pos_weight)
    names = list(p.keys())
    grads = torch.autograd.grad(loss, [p[n] for n in names], allow_unused=True)
    g, new = OrderedDict(), OrderedDict()
    for n, gr in zip(names, grads):
        gr = torch.zeros_like(p[n]) if gr is None else gr
        g[n] = gr.detach().numpy()
        new[n] = (p[n].detach() - torch.tensor(lr, dtype=dtype) * gr).numpy()
    return {"loss": float(loss.detach()), "loss_attention": float(la.detach()), "loss_classes": float(lc.detach()),
            "attentions": [a.detach().numpy() for a in att], "cls_logits": cls.detach().numpy(), "grads": g,
            "new_params": new}


# --------------------------------------------------------------------------
# F2: the top-level (current HEAD) LinkNet: vgg_16 trunk + deep-supervised U-shape (BAISNet.py:117-269) and its
# cal_loss (BAISRunnerTrain.py:97-115)
# --------------------------------------------------------------------------


def linknet_top_specs(width=1.0):
    specs = OrderedDict()

    def conv(name, k, cin, cout):
        specs[name + "/weights"] = (k, k, cin, cout)
        specs[name + "/biases"] = (cout,)

    cin, ch = 3, {}
    for blk, reps, c in VGG_BLOCKS:
        c = max(8, int(c * width))
        for r in range(1, reps + 1):
            conv("vgg_16/conv%d/conv%d_%d" % (blk, blk, r), 3, cin, c)
            cin = c
        ch[blk] = c
    c1, c2, c3, c4 = ch[2], ch[3], ch[4], ch[5]
    for lvl, c_in, c_out in ((4, c4, c3), (3, c3, c2), (2, c2, c1), (1, c1, c1)):
        for scope, has4 in (("attention_%d/segment_side_%d" % (lvl, lvl), True), ("attention_%d/%d" % (lvl, lvl), False)):
            conv(scope + "/d_s_conv_1", 1, c_in, c_in // 4)
            conv(scope + "/d_s_conv_2", 3, c_in // 4, c_in // 4)
            conv(scope + "/d_s_conv_3", 1, c_in // 4, c_out)
            if has4:
                conv(scope + "/d_s_conv_4", 3, c_out, 2)
    conv("attention_0", 3, c1, 2)
    return specs


def linknet_top_forward(params, image_nhwc):
    """LinkNet.build (BAISNet.py:169-269): returns the five segment logit maps [NHWC, 2 channels], coarse to fine."""
    def cv(x, n, stride=1, padding="SAME", relu=True):
        y = conv2d(x, params[n + "/weights"], stride, padding, bias=params[n + "/biases"])
        return F.relu(y) if relu else y

    x = image_nhwc.permute(0, 3, 1, 2)
    blocks = {}
    for blk, reps, _ in VGG_BLOCKS:
        for r in range(1, reps + 1):
            x = cv(x, "vgg_16/conv%d/conv%d_%d" % (blk, blk, r))
        blocks[blk] = x
        if blk < 5:
            x = F.max_pool2d(x, 2, 2)
    b1, b2, b3, b4 = blocks[2], blocks[3], blocks[4], blocks[5]

    def decoder(x, scope, out_hw, with_head):
        y = cv(x, scope + "/d_s_conv_1")
        y = resize_nearest(y, out_hw)
        y = cv(y, scope + "/d_s_conv_2")
        y = cv(y, scope + "/d_s_conv_3", padding="VALID")
        if with_head:
            return cv(y, scope + "/d_s_conv_4", padding="VALID", relu=False)
        return y

    segments = []
    cur = b4
    for lvl, nxt in ((4, b3), (3, b2), (2, b1), (1, b1)):
        if lvl == 1:
            cur = b1 + cur                                          # Net.add([block1, net_output]) (:244)
        hw = nxt.shape[2:4]
        segments.append(decoder(cur, "attention_%d/segment_side_%d" % (lvl, lvl), hw, True).permute(0, 2, 3, 1))
        cur = decoder(cur, "attention_%d/%d" % (lvl, lvl), hw, False)
    segments.append(cv(cur, "attention_0", padding="VALID", relu=False).permute(0, 2, 3, 1))
    return segments


def linknet_top_train_step(params_np, image, label_seg, lr=5e-3, dtype=torch.float64, pos_weight=1.0):
    p = to_torch(params_np, dtype, requires_grad=True)
    segs = linknet_top_forward(p, torch.as_tensor(np.asarray(image)).to(dtype))
    lab = torch.as_tensor(np.asarray(label_seg)).permute(0, 3, 1, 2).to(dtype)
    terms = []
    for a in segs:
        z = resize_nearest(lab, a.shape[1:3]).permute(0, 2, 3, 1).reshape(-1).long()
        terms.append(weighted_cross_entropy_with_logits(F.one_hot(z, 2).to(dtype), a.reshape(-1, 2), pos_weight).mean())
    loss = sum(terms) / len(terms)
    names = list(p.keys())
    grads = torch.autograd.grad(loss, [p[n] for n in names], allow_unused=True)
    g, new = OrderedDict(), OrderedDict()
    for n, gr in zip(names, grads):
        gr = torch.zeros_like(p[n]) if gr is None else gr
        g[n] = gr.detach().numpy()
        new[n] = (p[n].detach() - torch.tensor(lr, dtype=dtype) * gr).numpy()
    return {"loss": float(loss.detach()), "loss_segments": [float(t.detach()) for t in terms],
            "segments": [a.detach().numpy() for a in segs], "grads": g, "new_params": new}


# --------------------------------------------------------------------------
# Storage-rounding model of a 16-bit tensor-core path (test infrastructure for the bf16 parity bars)
# --------------------------------------------------------------------------

ROUNDING_POLICIES = {
    # what the CUDA bf16 path rounds to bf16 (everything else is float32): see pspnet_forward_rounded
    "round1": frozenset({"w", "conv", "act", "junc", "inc"}),           # round 1: every stored tensor
    "fused": frozenset({"w", "act", "junc", "inc"}),                    # BN applied from the fp32 accumulators
    "fused_fp32_trunk": frozenset({"w", "act", "junc_op", "inc"}),      # ... and the shortcut kept in float32
    "operands_only": frozenset({"w", "act", "junc_op"}),                # the floor of any bf16-operand MMA path
    "weights_only": frozenset({"w"}),
    "none": frozenset(),
}


def pspnet_forward_rounded(params, data_nhwc, last_pool_size, policy, seg_name="conv6_n",
                           r16_dtype=torch.bfloat16):
    """The forward pass of ``pspnet_forward`` (segment path) in the dtype of ``params`` with 16-bit roundings
    injected where a 16-bit tensor-core implementation stores or feeds 16-bit values.  It answers "how far from the
    float64 result does ANY implementation with this storage policy land" -- the bar the CUDA bf16 path is held to
    (tests/test_gpu_net.py), and the evidence for which roundings are irreducible (profiles/parity_r02.md).

    policy flags:  'w' bf16 weights of the tensor-core convolutions;  'conv' conv outputs rounded before BN;
    'act' BN-apply / pool / bilinear outputs rounded;  'junc' junction outputs (residual trunk) rounded;
    'junc_op' only the convolution-operand copy of the trunk rounded, the shortcut stays float32;
    'inc' the increase / projection conv outputs feeding a junction rounded.
    Returns {layer name: NCHW tensor} for the stage outputs and 'logits'."""
    def r16(t):
        return t.to(r16_dtype).to(t.dtype)

    ident = (lambda t: t)
    rw = r16 if "w" in policy else ident
    rc = r16 if "conv" in policy else ident
    ra = r16 if "act" in policy else ident
    ri = r16 if "inc" in policy else ident
    L = {}

    def W(n):
        return params[n + "/weights"]

    def BN(x, n, relu):
        return batch_norm(x, params["%s/%s/gamma" % (n, n)], params["%s/%s/beta" % (n, n)], relu)

    x = data_nhwc.permute(0, 3, 1, 2)
    x = ra(F.relu(BN(rc(conv2d(x, W("conv1_1_3x3_s2_n"), 2, "SAME")), "conv1_1_3x3_s2_bn", False)))   # fp32 stem
    x = ra(BN(rc(conv2d(x, rw(W("conv1_2_3x3")), 1, "SAME")), "conv1_2_3x3_bn", True))
    x = ra(BN(rc(conv2d(x, rw(W("conv1_3_3x3")), 1, "SAME")), "conv1_3_3x3_bn", True))
    L["conv1_3_3x3_bn"] = x
    x = max_pool_3x3_s2_same(x)
    x_sc = x
    for stage, blocks, _, stride, dil in STAGES:
        for b in range(1, blocks + 1):
            p = "conv%d_%d" % (stage, b)
            s = stride if b == 1 else 1
            if b == 1:
                sc = BN(ri(conv2d(x, rw(W(p + "_1x1_proj")), s)), p + "_1x1_proj_bn", False)
            else:
                sc = x_sc
            y = ra(BN(rc(conv2d(x, rw(W(p + "_1x1_reduce")), s)), p + "_1x1_reduce_bn", True))
            y = ra(BN(rc(conv2d(y, rw(W(p + "_3x3")), 1, dil, dil)), p + "_3x3_bn", True))
            y = BN(ri(conv2d(y, rw(W(p + "_1x1_increase")), 1)), p + "_1x1_increase_bn", False)
            t = F.relu(sc + y)
            if "junc" in policy:
                x = x_sc = r16(t)
            elif "junc_op" in policy:
                x, x_sc = r16(t), t
            else:
                x = x_sc = t
            L[p + "/relu"] = x
    c53 = x
    size = c53.shape[2:4]
    br = {}
    for lvl in PSP_LEVELS:
        n = "conv5_3_pool%d" % lvl
        y = ra(avg_pool(c53, last_pool_size // lvl))
        y = ra(BN(rc(conv2d(y, rw(W(n + "_conv")), 1)), n + "_conv_bn", True))
        br[lvl] = ra(resize_bilinear_ac(y, size))
        L[n + "_interp"] = br[lvl]
    cat = torch.cat([c53, br[6], br[3], br[2], br[1]], dim=1)
    y = ra(BN(rc(conv2d(cat, rw(W("conv5_4")), 1, "SAME")), "conv5_4_bn", True))
    L["conv5_4_bn"] = y
    L["logits"] = conv2d(y, W(seg_name), 1, bias=params[seg_name + "/biases"])      # fp32 head
    return L


# --------------------------------------------------------------------------
# F4: cascaded attention re-decoding (back/8AttentionU/BAISNet.py:485-650, BAISRunnerTrain.py:56-78,161-193):
# the 2AddClass trunk, then FOUR pyramid decoders.  Decoder k reads the feature gated by the softmax attention channel
# of decoder k-1 and every gate also feeds a class head; cal_loss scores the SIGMOID outputs (softmax CE on the two
# 4-channel ones, pos_weight-3 weighted BCE x 2 on channel 1 of the two 2-channel ones) and averages.
# --------------------------------------------------------------------------

CASCADE_SCOPES = ("", "attention_1/", "attention_2/", "attention_3/")


def attention_u_specs(num_classes=21, num_segment=4, filter_number=32, attention_module_num=2):
    """Ordered {tf variable name: shape}: trunk (2AddClass names), decoders in scopes '', attention_1..3 (the last
    `attention_module_num` with 2 output channels), class heads in attention_1..3 and at the top level."""
    base = param_specs("4BorderClass", num_classes, num_segment, filter_number)
    specs = OrderedDict((k, v) for k, v in base.items()
                        if not (k.startswith("conv5_3_pool") or k.startswith("conv5_4") or k.startswith("conv6_n_4")
                                or k.startswith("class_attention")))
    Fn = filter_number
    cin, psp = 32 * Fn, 8 * Fn

    def bn(name, c):
        leaf = name.split("/")[-1]
        specs["%s/%s/gamma" % (name, leaf)] = (c,)
        specs["%s/%s/beta" % (name, leaf)] = (c,)

    def decoder(sc, nseg):
        for lvl in PSP_LEVELS:
            specs[sc + "conv5_3_pool%d_conv/weights" % lvl] = (1, 1, cin, psp)
            bn(sc + "conv5_3_pool%d_conv_bn" % lvl, psp)
        specs[sc + "conv5_4/weights"] = (3, 3, 2 * cin, psp)
        bn(sc + "conv5_4_bn", psp)
        specs[sc + "conv6_n_4/weights"] = (1, 1, psp, nseg)
        specs[sc + "conv6_n_4/biases"] = (nseg,)

    def classifier(sc):
        specs[sc + "class_attention_conv/weights"] = (5, 5, cin, 16 * Fn)
        specs[sc + "class_attention_conv/biases"] = (16 * Fn,)
        specs[sc + "class_attention_fc/weights"] = (16 * Fn, num_classes)
        specs[sc + "class_attention_fc/biases"] = (num_classes,)

    n = len(CASCADE_SCOPES)
    for i, sc in enumerate(CASCADE_SCOPES):
        if i > 0:
            classifier(sc)             # (inside the scope the classifier comes before the decoder, :593-599)
        decoder(sc, 2 if i >= n - attention_module_num else num_segment)
    classifier("")                     # the last gate's class head sits outside any scope (:640-646)
    return specs


def attention_u_forward(params, data_nhwc, last_pool_size=40, num_segment=4, segment_attention=1,
                        attention_module_num=2):
    """-> (segments, attentions, classes): sigmoid outputs NHWC, gates N1HW, class logits, in build() order."""
    c53 = _pspnet_trunk(params, data_nhwc.permute(0, 3, 1, 2), {}, wiring="8AttentionU")
    P = last_pool_size
    size = c53.shape[2:4]

    def BN(x, n, relu):
        leaf = n.split("/")[-1]
        return batch_norm(x, params["%s/%s/gamma" % (n, leaf)], params["%s/%s/beta" % (n, leaf)], relu)

    def decoder(feat, sc):
        br = {}
        for lvl in PSP_LEVELS:
            n = sc + "conv5_3_pool%d" % lvl
            y = avg_pool(feat, P // lvl)
            y = BN(conv2d(y, params[n + "_conv/weights"], 1), n + "_conv_bn", True)
            br[lvl] = resize_bilinear_ac(y, size)
        cat = torch.cat([feat, br[6], br[3], br[2], br[1]], dim=1)
        y = BN(conv2d(cat, params[sc + "conv5_4/weights"], 1, "SAME"), sc + "conv5_4_bn", True)
        lg = conv2d(y, params[sc + "conv6_n_4/weights"], 1, bias=params[sc + "conv6_n_4/biases"])
        return torch.sigmoid(lg), torch.softmax(lg, dim=1)

    def classifier(m, sc):
        m = avg_pool(m, P // 5)
        m = F.relu(conv2d(m, params[sc + "class_attention_conv/weights"], 5,
                          bias=params[sc + "class_attention_conv/biases"]))
        assert m.shape[2] == 1 and m.shape[3] == 1
        return m[:, :, 0, 0] @ params[sc + "class_attention_fc/weights"] + params[sc + "class_attention_fc/biases"]

    segments, attentions, classes = [], [], []
    feat = c53
    n = len(CASCADE_SCOPES)
    sig, soft = decoder(feat, "")
    segments.append(sig.permute(0, 2, 3, 1))
    for i in range(1, n + 1):
        sel = segment_attention if (i - 1) < n - attention_module_num else 1      # (:586,603 vs :620,637)
        att = soft[:, sel:sel + 1]
        attentions.append(att)
        feat = feat * att
        sc = CASCADE_SCOPES[i] if i < n else ""
        classes.append(classifier(feat, sc))
        if i < n:
            sig, soft = decoder(feat, sc)
            segments.append(sig.permute(0, 2, 3, 1))
    return segments, attentions, classes


def attention_u_losses(segments, classes, label_segment, label_attention, label_classes, num_segment=4,
                       attention_module_num=2, pos_weight=3.0):
    """cal_loss (back/8AttentionU/BAISRunnerTrain.py:161-193) -> (loss, loss_segment_all, loss_class_all)."""
    ls = label_segment.reshape(-1).long()
    la = label_attention.reshape(-1)
    terms = []
    for i, sg in enumerate(segments):
        if i < len(segments) - attention_module_num:
            terms.append(F.cross_entropy(sg.reshape(-1, num_segment), ls, reduction="mean"))
        else:
            x = sg[..., 1].reshape(-1)
            terms.append(weighted_cross_entropy_with_logits(la.to(x.dtype), x, pos_weight).mean() * 2)
    lcs = [F.cross_entropy(c, label_classes.long(), reduction="mean") for c in classes]
    lseg, lcls = sum(terms) / len(terms), sum(lcs) / len(lcs)
    return lseg + 0.1 * lcls, lseg, lcls


def attention_u_train_step(params_np, data_nhwc, label_segment, label_attention, label_classes, last_pool_size=40,
                           lr=5e-3, dtype=torch.float64, num_segment=4, segment_attention=1, attention_module_num=2):
    p = to_torch(params_np, dtype, requires_grad=True)
    segs, atts, cls = attention_u_forward(p, torch.as_tensor(np.asarray(data_nhwc)).to(dtype), last_pool_size,
                                          num_segment, segment_attention, attention_module_num)
    loss, lseg, lcls = attention_u_losses(segs, cls, torch.as_tensor(np.asarray(label_segment)),
                                          torch.as_tensor(np.asarray(label_attention)),
                                          torch.as_tensor(np.asarray(label_classes)), num_segment, attention_module_num)
    names = list(p.keys())
    grads = torch.autograd.grad(loss, [p[n] for n in names], allow_unused=True)
    g, new = OrderedDict(), OrderedDict()
    for n, gr in zip(names, grads):
        gr = torch.zeros_like(p[n]) if gr is None else gr
        g[n] = gr.detach().numpy()
        new[n] = (p[n].detach() - torch.tensor(lr, dtype=dtype) * gr).numpy()
    return {"loss": float(loss.detach()), "loss_segment": float(lseg.detach()), "loss_classes": float(lcls.detach()),
            "segments": [t.detach().numpy() for t in segs], "attentions": [t.detach().numpy() for t in atts],
            "classes": [t.detach().numpy() for t in cls], "grads": g, "new_params": new}


# --------------------------------------------------------------------------
# A15..A18: losses, SGD, predictions
# --------------------------------------------------------------------------


def losses(seg_logits_nhwc, cls_logits, label_seg, label_cls, variant="2AddClass", pos_weight=3.0,
           class_weight=0.2):
    """(loss, loss_segment, loss_classes) exactly as build_net composes them."""
    _, _, has_class, kind = VARIANTS[variant]
    nseg = seg_logits_nhwc.shape[-1]
    if kind == "bce":
        pred = seg_logits_nhwc.reshape(-1)
        lab = label_seg.reshape(-1).to(pred.dtype)
        loss_seg = weighted_cross_entropy_with_logits(lab, pred, pos_weight).mean()
    else:
        pred = seg_logits_nhwc.reshape(-1, nseg)
        loss_seg = F.cross_entropy(pred, label_seg.reshape(-1).long(), reduction="mean")
    if has_class:
        loss_cls = F.cross_entropy(cls_logits, label_cls.long(), reduction="mean")
        return loss_seg + class_weight * loss_cls, loss_seg, loss_cls
    return loss_seg, loss_seg, torch.zeros((), dtype=loss_seg.dtype)


def poly_lr(base_lr, step, num_steps, power=0.9):
    """float32 arithmetic like the TF graph (step fed as float32)."""
    s = np.float32(step) / np.float32(num_steps)
    return np.float32(base_lr) * np.power(np.float32(1) - s, np.float32(power))


def to_torch(params, dtype=torch.float32, requires_grad=False):
    out = OrderedDict()
    for k, v in params.items():
        t = torch.tensor(np.asarray(v), dtype=dtype)
        t.requires_grad_(requires_grad)
        out[k] = t
    return out


def train_step(params_np, data_nhwc, label_seg, label_cls, variant="2AddClass", num_segment=1,
               last_pool_size=40, pos_weight=3.0, class_weight=0.2, lr=5e-3, dtype=torch.float32,
               attention_channel=None, keep=(), keep_grads=False):
    """One forward/backward/SGD step.  Returns dict(loss.., logits.., grads, new_params).

    keep: layer names whose activations are returned (NHWC); keep_grads additionally returns d loss / d layer."""
    seg_name, fc_name, has_class, _ = VARIANTS[variant]
    p = to_torch(params_np, dtype, requires_grad=True)
    x = torch.as_tensor(np.asarray(data_nhwc)).to(dtype)
    out = pspnet_forward(p, x, variant, num_segment, last_pool_size, attention_channel, keep)
    lab = torch.as_tensor(np.asarray(label_seg))
    cls = torch.as_tensor(np.asarray(label_cls)) if has_class else None
    loss, lseg, lcls = losses(out[seg_name], out.get(fc_name), lab, cls, variant, pos_weight, class_weight)
    names = list(p.keys())
    kept = [out["__nchw__"][k_] for k_ in keep] if keep_grads else []
    grads = torch.autograd.grad(loss, [p[n] for n in names] + kept, allow_unused=True)
    act_grads, grads = grads[len(names):], grads[:len(names)]
    g = OrderedDict()
    new = OrderedDict()
    for n, gr in zip(names, grads):
        gr = torch.zeros_like(p[n]) if gr is None else gr
        g[n] = gr.detach().numpy()
        new[n] = (p[n].detach() - torch.tensor(lr, dtype=dtype) * gr).numpy()
    res = {"loss": float(loss.detach()), "loss_segment": float(lseg.detach()), "loss_classes": float(lcls.detach()),
           "seg_logits": out[seg_name].detach().numpy(), "grads": g, "new_params": new}
    if has_class:
        res["cls_logits"] = out[fc_name].detach().numpy()
    for i, k_ in enumerate(keep):
        res[k_] = out[k_].detach().numpy()
        if keep_grads:
            g_ = act_grads[i]
            res["d:" + k_] = None if g_ is None else (g_.permute(0, 2, 3, 1) if g_.dim() == 4 else g_).detach().numpy()
    return res


def predict_train(seg_logits, cls_logits=None):
    """Training-time predictions: binary head thresholds the *logits* at 0.5; multi-class argmax."""
    if seg_logits.shape[-1] == 1:
        pred = (seg_logits > 0.5).astype(np.int32)
    else:
        pred = np.argmax(seg_logits, axis=-1).astype(np.int32)[..., None]
    cls = None if cls_logits is None else np.argmax(cls_logits, axis=-1).astype(np.int32)
    return pred, cls


def predict_click(seg_logits_nhwc, input_size=None):
    """Runner / RunnerGUI mask: argmax(sigmoid(logits)) (optionally after the legacy upsample)."""
    x = torch.as_tensor(np.asarray(seg_logits_nhwc), dtype=torch.float32).permute(0, 3, 1, 2)
    if input_size is not None:
        x = resize_bilinear_legacy(x, input_size)
    s = torch.sigmoid(x)
    return torch.argmax(s, dim=1).numpy().astype(np.int64)
