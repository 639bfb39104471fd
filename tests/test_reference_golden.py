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
