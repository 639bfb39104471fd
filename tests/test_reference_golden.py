"""The oracle, the host `Data` class and the CUDA kernels of the data rows (A1 click map, A2 pack, A3 label encodings,
F3 click sampling, cfg1 load_image) against outputs of the REFERENCE's own code.

tests/golden/reference_data.npz was produced by tests/golden/make_reference_golden.py, which imports
/root/reference/back/{2AddClass,3ThreeClass,4BorderClass,8AttentionU}/BAISData.py (numpy + PIL only, no TensorFlow) in
the build container and runs them on tests/golden/voc_mini/ and on the reference's fixture input/7.jpg.  Nothing here
reads /root/reference.  Everything is compared bit for bit."""
import ctypes as C
import os

import numpy as np
import pytest

from oracle import basi_oracle as O

HERE = os.path.dirname(os.path.abspath(__file__))
VOC = os.path.join(HERE, "golden", "voc_mini") + "/"
SIZE, RATIO = (64, 64), 8


@pytest.fixture(scope="module")
def G():
    return np.load(os.path.join(HERE, "golden", "reference_data.npz"))


def _bits(a):
    a = np.ascontiguousarray(a)
    return a.view(np.uint32) if a.dtype == np.float32 else a


def _n_cases(G, prefix):
    i = 0
    while "%s/%d/args" % (prefix, i) in G:
        i += 1
    return i


# ------------------------------------------------------------------ CPU: oracle + host mirror vs the reference
def test_oracle_click_map_equals_reference_bitwise(G):
    from basi_b200.BAISData import Data, click_lut
    n = _n_cases(G, "mask_gaussian")
    assert n >= 5
    for i in range(n):
        h, w, y0, x0, sigma = [int(v) for v in G["mask_gaussian/%d/args" % i]]
        ref = G["mask_gaussian/%d/out" % i]
        assert ref.dtype == np.float32 and ref.shape == (h, w)
        assert np.array_equal(_bits(O.mask_gaussian((h, w), (y0, x0), sigma)), _bits(ref))
        assert np.array_equal(_bits(Data._mask_gaussian((h, w), [y0, x0], sigma)), _bits(ref))
        # the device kernel gathers this table by integer squared distance
        lut = click_lut((h, w), sigma)
        yy, xx = np.mgrid[0:h, 0:w]
        assert np.array_equal(_bits(lut[(yy - y0) ** 2 + (xx - x0) ** 2]), _bits(ref))


@pytest.mark.parametrize("tag,fn", [("2AddClass", "encode_labels_binary"), ("3ThreeClass", "encode_labels_three"),
                                    ("4BorderClass_plain", "encode_labels_three"),
                                    ("4BorderClass_255", "encode_labels_border"),
                                    ("8AttentionU", "encode_labels_border")])
def test_oracle_label_encodings_equal_reference(G, tag, fn):
    """_read_annotation of each snapshot (binary, (x-1)//127 three-class, has_255 four-class with its uint8 wrap)."""
    names = ["a", "b", "c"]
    idx, nums, masks = G[tag + "/ann_index"], G[tag + "/ann_num"], G[tag + "/ann_mask"]
    assert len(idx) == int(G[tag + "/n_ann"]) >= 7
    for i in range(len(idx)):
        raw = G["raw/%s/obj" % names[int(idx[i])]]
        assert raw.dtype == np.uint8
        got = getattr(O, fn)(raw, int(nums[i]))
        assert np.array_equal(np.asarray(got).astype(np.int64), masks[i]), (tag, i)
    if fn == "encode_labels_border":
        assert set(np.unique(masks)) == {1, 2, 3} or set(np.unique(masks)) == {0, 1, 2, 3}
        assert np.all(masks[0][G["raw/a/obj"] == 255] == 2) and np.all(masks[0][G["raw/a/obj"] == 0] == 3)


def test_host_data_class_equals_reference_2addclass(G):
    """basi_b200.BAISData.Data on the same files with the same numpy RNG state: annotations, class ids (incl. the
    out-of-range id -> 0 rule), decoded images, sampled clicks, click maps and packed batches, three batches
    (the third crosses an epoch and reshuffles)."""
    from basi_b200.BAISData import Data
    tag = "2AddClass"
    d = Data(data_list="ImageSets/Segmentation/train.txt", data_path="JPEGImages/", data_root_path=VOC,
             annotation_path="SegmentationObject/", class_path="SegmentationClass/", batch_size=2, image_size=SIZE,
             ratio=RATIO)
    assert len(d._annotations) == int(G[tag + "/n_ann"])
    assert [a[0] for a in d._annotations] == list(G[tag + "/ann_index"])
    assert [a[1] for a in d._annotations] == list(G[tag + "/ann_num"])
    assert [a[2] for a in d._annotations] == list(G[tag + "/ann_class"])
    assert 0 in list(G[tag + "/ann_class"])                                   # the class id 200 of image c
    assert np.array_equal(np.stack([a[3] for a in d._annotations]), G[tag + "/ann_mask"])
    assert np.array_equal(_bits(np.stack(d._images_data)), _bits(G[tag + "/images"]))
    np.random.seed(11)
    for step in range(3):
        data, ann, cls, _, mask = d.next_batch_train()
        assert np.array_equal(_bits(np.stack(data)), _bits(G["%s/step%d/data" % (tag, step)])), step
        assert np.array_equal(np.stack(ann), G["%s/step%d/ann" % (tag, step)])
        assert list(cls) == list(G["%s/step%d/cls" % (tag, step)])
        assert np.array_equal(_bits(np.stack(mask)), _bits(G["%s/step%d/mask" % (tag, step)]))


def test_load_image_equals_reference_on_its_fixture(G):
    from basi_b200.BAISData import Data
    path = os.path.join(HERE, "golden", "input_7.jpg")
    final, raw, _ = Data.load_image(path, where=[40, 25], image_size=(64, 64))[:3]
    assert np.array_equal(_bits(final[0]), _bits(G["load_image/file/final"]))
    assert np.array_equal(_bits(raw), _bits(G["load_image/file/raw"]))
    from PIL import Image
    final2 = Data.load_image(np.asarray(Image.open(path)), where=[12, 60], image_size=(64, 64))[0]
    assert np.array_equal(_bits(final2[0]), _bits(G["load_image/array/final"]))
    # the oracle's pack (uint8 image + click) gives the same 4-channel input
    img_u8 = G["load_image/file/raw"].astype(np.uint8)
    assert np.array_equal(_bits(O.pack_input(img_u8, (40, 25))), _bits(G["load_image/file/final"]))


def test_cascade_attention_labels_equal_reference(G):
    for step in range(3):
        ann, att = G["8AttentionU/step%d/ann" % step], G["8AttentionU/step%d/ann_attention" % step]
        assert np.array_equal((ann == 1).astype(np.int64), att)


# ------------------------------------------------------------------ GPU: the kernels vs the reference
@pytest.mark.gpu
def test_clickmap_pack_kernel_equals_reference_batches(G):
    """basi_clickmap_pack (uint8 image + click -> NHWC4 float32) == the reference's np.concatenate((image / 255,
    _mask_gaussian(click)), 2) for the batches its own next_batch_train produced."""
    import torch
    from gpu_util import call, dev, host
    from basi_b200.BAISData import click_lut
    lut = click_lut(SIZE, 30)
    lutd = dev(lut)
    for tag in ("2AddClass", "4BorderClass_255"):
        for step in range(3):
            ref = G["%s/step%d/data" % (tag, step)]                      # [B, 64, 64, 4]
            B = ref.shape[0]
            img = np.rint(ref[..., :3] * 255).astype(np.uint8)
            assert np.array_equal(_bits(img.astype(np.float32) / 255), _bits(ref[..., :3]))
            m = ref[..., 3]
            clicks = np.asarray([np.unravel_index(np.argmax(m[b]), m[b].shape) for b in range(B)], dtype=np.int32)
            out = torch.zeros((B, 64, 64, 4), dtype=torch.float32, device="cuda:0")
            imgd, clickd = dev(img), dev(clicks)
            call("basi_clickmap_pack", imgd.data_ptr(), 0, clickd.data_ptr(), lutd.data_ptr(), C.c_int64(lut.size),
                 out.data_ptr(), B, 64, 64)
            assert np.array_equal(_bits(host(out)), _bits(ref)), (tag, step)


@pytest.mark.gpu
@pytest.mark.parametrize("tag,mode", [("2AddClass", 0), ("3ThreeClass", 2), ("4BorderClass_255", 1), ("8AttentionU", 1)])
def test_label_encode_kernel_equals_reference(G, tag, mode):
    import torch
    from gpu_util import call, dev, host
    names = ["a", "b", "c"]
    idx, nums, masks = G[tag + "/ann_index"], G[tag + "/ann_num"], G[tag + "/ann_mask"]
    B = len(idx)
    ann = np.stack([G["raw/%s/obj" % names[int(i)]] for i in idx]).astype(np.uint8)
    P = ann.shape[1]
    oi = torch.full((B, P, P), -7, dtype=torch.int32, device="cuda:0")
    of = torch.full((B, P, P), -7.0, dtype=torch.float32, device="cuda:0")
    annd, numd = dev(ann), dev(nums.astype(np.int32))
    call("basi_label_encode", annd.data_ptr(), None, numd.data_ptr(), mode, oi.data_ptr(), of.data_ptr(), B,
         C.c_int64(P * P))
    assert np.array_equal(host(oi).astype(np.int64), masks)
    assert np.array_equal(host(of).astype(np.int64), masks)


@pytest.mark.gpu
def test_engine_annotation_feed_reproduces_reference_batch(G):
    """Engine.feed_annotations + the click-map kernel with the reference's RNG state: 4-class labels, sampled clicks
    and the packed network input of the reference's first 4BorderClass(has_255) batch, bit for bit."""
    import torch
    from basi_b200.BAISPSPNet import PSPNet, Placeholder
    from basi_b200.engine import Engine
    tag = "4BorderClass_255"
    names = ["a", "b", "c"]
    idx, nums = G[tag + "/ann_index"], G[tag + "/ann_num"]
    B, S, P = 2, 64, 8
    net = PSPNet({'data': Placeholder((None, S, S, 4))}, num_classes=21, num_segment=4, is_training=True,
                 last_pool_size=P, filter_number=8, variant="4BorderClass")
    eng = Engine(net, B, "f32", True, dict(kind="softmax", class_weight=0.1))
    eng.enable_click_input(30)
    eng.enable_label_input("border")
    np.random.seed(11)                       # the state the reference's first next_batch_train started from
    sel = [0, 1]                             # its first batch: annotations 0 and 1 (no shuffle before the first epoch)
    ann = np.stack([G["raw/%s/obj" % names[int(idx[i])]] for i in sel]).astype(np.uint8)
    eng.feed_annotations(ann, nums[sel].astype(np.int32))
    ref = G[tag + "/step0/data"]
    img = np.rint(ref[..., :3] * 255).astype(np.uint8)
    eng.img_u8.copy_(torch.from_numpy(img).to(eng.img_u8.device).view(eng.img_u8.shape))
    eng._run(eng.pre, eng._stream())
    torch.cuda.synchronize()
    assert np.array_equal(eng.label_seg.cpu().numpy().astype(np.int64), G[tag + "/step0/ann"])
    assert np.array_equal(_bits(eng.input.t.cpu().numpy()), _bits(ref))


# ===============================================================================================================
# Golden vectors of the reference's own (vendored TF-slim) tests for ops on the hot path:
# tests/golden/slim_reference_tests.json, extracted by tests/golden/make_slim_golden.py from
# slim/nets/resnet_v1_test.py:58-153 (ResnetUtilsTest: subsample, conv2d 'SAME' stride 1 / 2 on even and odd inputs,
# conv2d_same = explicit padding + 'VALID') and slim/nets/vgg_test.py:230-333 (vgg_16 end points / variable names).
# They pin the TF padding semantics of Network.conv (back/2AddClass/BAISPSPNet.py:118-146) -- in particular the
# asymmetric (0 before, 1 after) 'SAME' padding of conv1_1_3x3_s2 on an even input -- and the strided 1x1 convolution
# (= subsample) in the oracle, in the host lowering rule and in the CUDA kernels.  All values are small integers:
# exact in float32, bfloat16 and fp16, so every comparison is bit for bit.
# ===============================================================================================================
@pytest.fixture(scope="module")
def SLIM():
    import json
    with open(os.path.join(HERE, "golden", "slim_reference_tests.json")) as f:
        return json.load(f)


def _mesh(n):
    """create_test_input(1, n, n, 1): x[h, w] = h + w (slim/nets/resnet_v1_test.py:30-53)"""
    return (np.arange(n).reshape(n, 1) + np.arange(n).reshape(1, n)).astype(np.float32)


@pytest.mark.parametrize("case", ["conv2d_same_even", "conv2d_same_odd"])
def test_oracle_conv_padding_semantics_equal_slim_golden(SLIM, case):
    import torch
    g = SLIM["resnet_utils"][case]
    n = g["n"]
    x = torch.from_numpy(_mesh(n)).view(1, 1, n, n)
    w = torch.from_numpy(_mesh(3)).view(3, 3, 1, 1)                       # HWIO
    want = {k: np.asarray(g[k], np.float32) for k in ("y1_expected", "y2_expected", "y3_expected", "y4_expected")}
    y1 = O.conv2d(x, w, 1, "SAME")
    assert np.array_equal(y1[0, 0].numpy(), want["y1_expected"])
    # resnet_utils.subsample(y1, 2) == the strided 1x1 'VALID' convolution of the trunk with a unit weight
    y2 = O.conv2d(y1, torch.ones(1, 1, 1, 1), 2, "VALID")
    assert np.array_equal(y2[0, 0].numpy(), want["y2_expected"])
    assert np.array_equal(y1[0, 0, ::2, ::2].numpy(), want["y2_expected"])
    # conv2d_same(stride 2) = pad (k-1)//2 on both sides, then 'VALID': Network.zero_padding(1) + conv(..., 'VALID')
    y3 = O.conv2d(x, w, 2, 1)
    assert np.array_equal(y3[0, 0].numpy(), want["y3_expected"])
    # tf 'SAME' with stride 2: (0, 1) padding on the even input, (1, 1) on the odd one
    y4 = O.conv2d(x, w, 2, "SAME")
    assert np.array_equal(y4[0, 0].numpy(), want["y4_expected"])
    if n % 2 == 0:
        assert not np.array_equal(want["y4_expected"], want["y2_expected"])       # the trap is real on even inputs
        assert O.tf_same_pad(n, 3, 2) == (0, 1)
    else:
        assert O.tf_same_pad(n, 3, 2) == (1, 1)


def test_oracle_subsample_equals_slim_golden(SLIM):
    for key in ("subsample_3x3", "subsample_4x4"):
        g = SLIM["resnet_utils"][key]
        x = np.arange(g["range"], dtype=np.float32).reshape(g["shape"])
        got = x[:, ::g["factor"], ::g["factor"], :]
        assert got.shape == tuple(g["expected_shape"])
        assert np.array_equal(got.reshape(-1), np.asarray(g["expected"], np.float32))


def test_host_same_padding_rule_equals_oracle_and_slim_golden(SLIM):
    """engine._same_pad_before is what the lowering writes into basi_conv_desc.pad_t / pad_l"""
    from basi_b200.engine import _same_pad_before
    assert _same_pad_before(SLIM["resnet_utils"]["conv2d_same_even"]["n"], 3, 2, 1) == 0
    assert _same_pad_before(SLIM["resnet_utils"]["conv2d_same_odd"]["n"], 3, 2, 1) == 1
    for n in range(1, 70):
        for k in (1, 2, 3, 5, 7):
            for s in (1, 2, 3, 5):
                for d in (1, 2, 4):
                    assert _same_pad_before(n, k, s, d) == O.tf_same_pad(n, k, s, d)[0], (n, k, s, d)


def test_variant_b_trunk_names_equal_slim_vgg16_test_lists(SLIM):
    """The vgg_16 trunk of variant B (slim/nets/vgg.py:187-196 up to conv5_3) creates exactly the convolution
    variables and end points slim's own vgg_16 tests expect (everything before pool5 / fc6)."""
    from basi_b200.BAISNet import LinkNet
    from basi_b200.BAISPSPNet import Placeholder
    net = LinkNet(Placeholder((None, 64, 64, 3)), Placeholder((None, 64, 64, 1)), True, 21)
    net.build()
    want_vars = [v for v in SLIM["vgg_16"]["model_variables"] if "/fc" not in v]
    got_vars = [v for v in net.variables if v.startswith("vgg_16/")]
    assert len(want_vars) == 26 and set(got_vars) == set(want_vars)
    want_eps = [e for e in SLIM["vgg_16"]["end_points"] if "/fc" not in e and not e.endswith("pool5")]
    got_eps = [e for e in net.layers if e.startswith("vgg_16/")]
    assert set(got_eps) == set(want_eps), set(got_eps) ^ set(want_eps)
    # shapes: five stride-2 stages minus pool5 -> conv5_3 at 1/16 resolution, slim channel widths
    assert tuple(net.layers["vgg_16/conv5/conv5_3"].shape) == (4, 4, 512)
    assert tuple(net.layers["vgg_16/pool1"].shape) == (32, 32, 64)


# ------------------------------------------------------------------ GPU: the CUDA kernels against the same vectors
def _embed(a_hw, c, n=1):
    """[H, W] -> NHWC with the values in channel 0 and zeros elsewhere"""
    out = np.zeros((n,) + a_hw.shape + (c,), np.float32)
    out[..., 0] = a_hw
    return out


@pytest.mark.gpu
@pytest.mark.parametrize("dtype", ["f32", "h16"])
@pytest.mark.parametrize("case", ["conv2d_same_even", "conv2d_same_odd"])
def test_cuda_conv_kernels_equal_slim_golden(SLIM, case, dtype):
    """basi_conv_fprop (the CUDA-core path every shape can take) for y1 / y3 / y4, basi_subsample_fwd for y2 and the
    conv1_1 stem kernel (basi_stem_fprop_stats: 4 -> 32 channels, 'SAME' stride 2) for y4, through the C ABI."""
    import torch
    from gpu_util import act, call, dev, empty_act, host
    from basi_b200 import _lib
    from basi_b200._lib import ConvDesc
    g = SLIM["resnet_utils"][case]
    n, n2 = g["n"], g["n2"]
    want = {k: np.asarray(g[k], np.float32) for k in ("y1_expected", "y2_expected", "y3_expected", "y4_expected")}
    td = torch.float32 if dtype == "f32" else torch.bfloat16        # (dtype code 1 = the build's 16-bit type)
    cin, cout = 8, 8
    x = _embed(_mesh(n), cin)
    w = np.zeros((3, 3, cin, cout), np.float32)
    w[:, :, 0, 0] = _mesh(3)
    xa, wd = act(x, td), dev(w)

    def conv(stride, pt, pl, out_n):
        ya = empty_act((1, out_n, out_n, cout), td, fill=7.0)
        desc = ConvDesc(3, 3, stride, 1, pt, pl, 0)
        call("basi_conv_fprop", C.byref(desc), xa.ref, wd.data_ptr(), None, ya.ref)
        y = host(ya)
        assert not np.any(y[..., 1:])                               # the other output channels have zero weights
        return y[0, :, :, 0], ya

    y1, y1a = conv(1, 1, 1, n)
    assert np.array_equal(y1, want["y1_expected"])
    y2a = empty_act((1, n2, n2, cout), td, fill=7.0)
    call("basi_subsample_fwd", y1a.ref, 2, y2a.ref)
    assert np.array_equal(host(y2a)[0, :, :, 0], want["y2_expected"])
    y3, _ = conv(2, 1, 1, n2)                                       # explicit padding 1, then 'VALID'
    assert np.array_equal(y3, want["y3_expected"])
    pad = O.tf_same_pad(n, 3, 2)[0]
    y4, _ = conv(2, pad, pad, n2)                                   # tf 'SAME', stride 2
    assert np.array_equal(y4, want["y4_expected"])
    # the stem kernel: float32 NHWC4 input, 32 output channels, fused batch-norm statistics
    x4 = act(_embed(_mesh(n), 4))
    w4 = np.zeros((3, 3, 4, 32), np.float32)
    w4[:, :, 0, 0] = _mesh(3)
    ys = empty_act((1, n2, n2, 32), td, fill=7.0)
    desc = ConvDesc(3, 3, 2, 1, pad, pad, 0)
    if _lib.load().basi_stem_fprop_stats_supported(C.byref(desc), x4.ref, ys.ref) == 1:
        sums = torch.zeros(2 * 32 * 8, dtype=torch.float64, device="cuda:0")
        bnp = torch.zeros(4 * 32, device="cuda:0")
        cnt = torch.zeros(2, dtype=torch.int32, device="cuda:0")
        gd, bd = dev(np.ones(32, np.float32)), dev(np.zeros(32, np.float32))
        call("basi_stem_fprop_stats", C.byref(desc), x4.ref, dev(w4).data_ptr(), ys.ref, sums.data_ptr(), gd.data_ptr(),
             bd.data_ptr(), C.c_double(n2 * n2), C.c_float(1e-5), bnp.data_ptr(), cnt.data_ptr())
        assert np.array_equal(host(ys)[0, :, :, 0], want["y4_expected"])
        assert abs(float(host(bnp)[0]) - float(want["y4_expected"].mean())) < 1e-4      # fused mean of channel 0


@pytest.mark.gpu
@pytest.mark.parametrize("case", ["conv2d_same_even", "conv2d_same_odd"])
def test_tcgen05_conv_equals_slim_golden(SLIM, case):
    """The tensor-core implicit-GEMM plan (TMA zero fill is the padding) on the same vectors: 3x3 stride 1 'SAME' with
    the golden image in channel 0 of a 64-channel tensor.  Small integers: exact in the 16-bit operands."""
    import torch
    from gpu_util import act, call, empty_act, host, stream
    from basi_b200 import _lib
    from basi_b200._lib import ConvDesc
    g = SLIM["resnet_utils"][case]
    n = g["n"]
    cin = cout = 64
    x = _embed(_mesh(n), cin)
    w = np.zeros((3, 3, cin, cout), np.float32)
    w[:, :, 0, 0] = _mesh(3)
    xa, ya = act(x, torch.bfloat16), empty_act((1, n, n, cout), torch.bfloat16, fill=7.0)
    desc = ConvDesc(3, 3, 1, 1, 1, 1, 0)
    lib = _lib.load()
    if lib.basi_tc_conv_supported(_lib.TC_FPROP, C.byref(desc), xa.ref, ya.ref) != 1:
        pytest.skip("shape not on the tcgen05 path")
    wt = torch.from_numpy(w).to("cuda:0")
    w_oi = torch.zeros(9 * cout * cin, dtype=torch.bfloat16, device="cuda:0")
    w_io = torch.zeros(9 * cin * cout, dtype=torch.bfloat16, device="cuda:0")
    call("basi_tc_pack_weights", wt.data_ptr(), w_io.data_ptr(), w_oi.data_ptr(), 9, cin, cout)
    h = C.c_void_p()
    _lib.call("basi_tc_conv_create", _lib.TC_FPROP, C.byref(desc), xa.ref, ya.ref, w_oi.data_ptr(), None, 0, C.byref(h))
    try:
        assert lib.basi_tc_conv_run(h, stream()) == 0
        y = host(ya)
    finally:
        lib.basi_tc_conv_destroy(h)
    assert np.array_equal(y[0, :, :, 0], np.asarray(g["y1_expected"], np.float32))
    assert not np.any(y[..., 1:])


# ------------------------------------------------------------------ the reference's own PROPERTY test of the atrous path
# slim/nets/resnet_v1_test.py:197-244 (testAtrousValuesBottleneck, 30 x 31 input: one odd and one even dimension):
# "dense feature extraction by atrous convolution followed by subsampling gives identical results to feature
# extraction at the nominal stride".  Restated on the primitives of this path (A5: zero_padding(d) + atrous conv d,
# back/2AddClass/BAISPSPNet.py:135-146): subsample(atrous_d(x), d) == conv_1(subsample(x, d)) with shared weights.
@pytest.mark.parametrize("rate", [2, 4])
def test_oracle_atrous_then_subsample_equals_nominal_stride(rate):
    import torch
    rng = np.random.RandomState(rate)
    x = torch.from_numpy(rng.uniform(-1, 1, (2, 6, 30, 31))).double()
    w = torch.from_numpy(rng.uniform(-1, 1, (3, 3, 6, 5))).double()
    dense = O.conv2d(x, w, 1, rate, rate)[:, :, ::rate, ::rate]
    nominal = O.conv2d(x[:, :, ::rate, ::rate], w, 1, 1, 1)
    assert dense.shape == nominal.shape
    assert float((dense - nominal).abs().max()) < 1e-12


@pytest.mark.gpu
@pytest.mark.parametrize("rate,ch,path", [(2, 16, "simt"), (4, 16, "simt"), (2, 128, "tc"), (4, 64, "tc")])
def test_cuda_atrous_then_subsample_equals_nominal_stride(rate, ch, path):
    """The same property through the C ABI: CUDA-core fp32 convolution and the tcgen05 plan (TMA zero fill = the
    zero_padding(d) of the reference).  Same products in the same k order for every output pixel: bit for bit."""
    import torch
    from gpu_util import act, bf16_round, call, dev, empty_act, host, stream
    from basi_b200 import _lib
    from basi_b200._lib import ConvDesc
    rng = np.random.RandomState(10 * rate + ch)
    B, H, W = 2, 30, 31
    x = rng.uniform(-1, 1, (B, H, W, ch)).astype(np.float32)
    w = (rng.uniform(-1, 1, (3, 3, ch, ch)) / np.sqrt(9 * ch)).astype(np.float32)
    td = torch.float32 if path == "simt" else torch.bfloat16
    if path == "tc":
        x, w = bf16_round(x), bf16_round(w)
    xs = np.ascontiguousarray(x[:, ::rate, ::rate, :])
    hs, ws = xs.shape[1], xs.shape[2]
    lib = _lib.load()

    def run(xin, d, oh, ow):
        xa, ya = act(xin, td), empty_act((B, oh, ow, ch), td, fill=7.0)
        desc = ConvDesc(3, 3, 1, d, d, d, 0)
        if path == "simt":
            call("basi_conv_fprop", C.byref(desc), xa.ref, dev(w).data_ptr(), None, ya.ref)
            return host(ya)
        if lib.basi_tc_conv_supported(_lib.TC_FPROP, C.byref(desc), xa.ref, ya.ref) != 1:
            pytest.skip("shape not on the tcgen05 path")
        w_oi = torch.zeros(9 * ch * ch, dtype=td, device="cuda:0")
        w_io = torch.zeros(9 * ch * ch, dtype=td, device="cuda:0")
        call("basi_tc_pack_weights", dev(w).data_ptr(), w_io.data_ptr(), w_oi.data_ptr(), 9, ch, ch)
        h = C.c_void_p()
        _lib.call("basi_tc_conv_create", _lib.TC_FPROP, C.byref(desc), xa.ref, ya.ref, w_oi.data_ptr(), None, 0,
                  C.byref(h))
        try:
            assert lib.basi_tc_conv_run(h, stream()) == 0
            return host(ya)
        finally:
            lib.basi_tc_conv_destroy(h)

    dense = run(x, rate, H, W)[:, ::rate, ::rate, :]
    nominal = run(xs, 1, hs, ws)
    assert dense.shape == nominal.shape
    assert np.array_equal(dense, nominal)
    # and both equal the oracle (float64) within the storage precision
    import torch as _t
    ref = O.conv2d(_t.from_numpy(xs).permute(0, 3, 1, 2).double(), _t.from_numpy(w).double(), 1, 1, 1)
    err = np.abs(nominal - ref.permute(0, 2, 3, 1).numpy()).max()
    assert err < (1e-5 if path == "simt" else 2e-2), err
