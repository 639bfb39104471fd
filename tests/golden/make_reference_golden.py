"""Golden vectors from the REFERENCE's own data code, run in the build container.

The host-side data path of the reference (BAISData.py of each snapshot: annotation decoding / label encodings,
click sampling, the Gaussian click map, image + click-map packing) is plain numpy + PIL and imports without
TensorFlow, so -- unlike the network arithmetic -- it can be executed here.  This script imports those modules
from /root/reference (read-only, never copied), runs them on a tiny synthetic VOC tree (tests/golden/voc_mini/,
written by this script and committed) and on the reference's own fixture input/7.jpg, and stores what they return
in tests/golden/reference_data.npz.  tests/test_reference_golden.py then holds the oracle, the host `Data` class and
the CUDA kernels to these outputs bit for bit.  /root/reference does not exist on the GPU box; only this script reads it.

    python tests/golden/make_reference_golden.py
"""
import importlib.util
import os
import sys

import numpy as np
from PIL import Image, ImageDraw

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference"
VOC = os.path.join(HERE, "voc_mini") + "/"
SIZE = (64, 64)          # square, like every reference run: image_size is used both as PIL (w, h) and numpy (rows, cols)
NAMES = ["a", "b", "c"]


def write_voc_mini():
    """3 images 120x90 with 2-3 instances each, a 255 border ring, class ids incl. one out-of-range value."""
    rng = np.random.RandomState(3)
    for d in ("ImageSets/Segmentation", "JPEGImages", "SegmentationObject", "SegmentationClass"):
        os.makedirs(VOC + d, exist_ok=True)
    with open(VOC + "ImageSets/Segmentation/train.txt", "w") as f:
        f.write("\n".join(NAMES) + "\n")
    pal = [0] * 768
    for i in range(256):
        pal[3 * i:3 * i + 3] = [(i * 37) % 256, (i * 91) % 256, (i * 13) % 256]
    for k, n in enumerate(NAMES):
        img = rng.randint(0, 256, size=(90, 120, 3), dtype=np.uint8)
        img = np.asarray(Image.fromarray(img).resize((120, 90), Image.BILINEAR))       # (smooth it a little)
        Image.fromarray(img).save(VOC + "JPEGImages/%s.jpg" % n, quality=92)
        obj = Image.new("P", (120, 90), 0)
        cls = Image.new("P", (120, 90), 0)
        obj.putpalette(pal)
        cls.putpalette(pal)
        do, dc = ImageDraw.Draw(obj), ImageDraw.Draw(cls)
        shapes = [((10, 10, 50, 45), 1, 15), ((60, 20, 110, 80), 2, 7 + k), ((20, 55, 55, 85), 3, 200 if k == 2 else 12)]
        for (box, inst, c) in shapes[: 2 + (k != 0)]:
            do.ellipse([box[0] - 2, box[1] - 2, box[2] + 2, box[3] + 2], fill=255)    # border ring
            dc.ellipse([box[0] - 2, box[1] - 2, box[2] + 2, box[3] + 2], fill=255)
            do.ellipse(box, fill=inst)
            dc.ellipse(box, fill=c)
        obj.save(VOC + "SegmentationObject/%s.png" % n)
        cls.save(VOC + "SegmentationClass/%s.png" % n)


def load_ref(snapshot):
    path = os.path.join(REF, "back", snapshot, "BAISData.py") if snapshot else os.path.join(REF, "BAISData.py")
    spec = importlib.util.spec_from_file_location("ref_BAISData_" + (snapshot or "top"), path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def run_snapshot(out, snapshot, tag, seed=11, **kw):
    mod = load_ref(snapshot)
    d = mod.Data(data_list="ImageSets/Segmentation/train.txt", data_path="JPEGImages/", data_root_path=VOC,
                 annotation_path="SegmentationObject/", class_path="SegmentationClass/", batch_size=2,
                 image_size=SIZE, ratio=8, **kw)
    out[tag + "/n_ann"] = np.int64(len(d._annotations))
    out[tag + "/ann_index"] = np.asarray([a[0] for a in d._annotations], np.int64)
    out[tag + "/ann_num"] = np.asarray([a[1] for a in d._annotations], np.int64)
    out[tag + "/ann_class"] = np.asarray([a[2] for a in d._annotations], np.int64)
    out[tag + "/ann_mask"] = np.stack([np.asarray(a[3]) for a in d._annotations]).astype(np.int64)
    out[tag + "/images"] = np.stack(d._images_data)
    np.random.seed(seed)
    for step in range(3):                      # the third call crosses number_patch and reshuffles
        r = d.next_batch_train()
        out["%s/step%d/data" % (tag, step)] = np.stack(r[0])
        out["%s/step%d/ann" % (tag, step)] = np.stack(r[1]).astype(np.int64)
        if len(r) == 6:                        # 8AttentionU: + attention labels
            out["%s/step%d/ann_attention" % (tag, step)] = np.stack(r[2]).astype(np.int64)
            out["%s/step%d/cls" % (tag, step)] = np.asarray(r[3], np.int64)
        else:
            out["%s/step%d/cls" % (tag, step)] = np.asarray(r[2], np.int64)
        out["%s/step%d/mask" % (tag, step)] = np.stack(r[-1])


def main():
    write_voc_mini()
    out = {}
    run_snapshot(out, "2AddClass", "2AddClass")
    run_snapshot(out, "3ThreeClass", "3ThreeClass")
    run_snapshot(out, "4BorderClass", "4BorderClass_plain", has_255=False)
    run_snapshot(out, "4BorderClass", "4BorderClass_255", has_255=True)
    run_snapshot(out, "8AttentionU", "8AttentionU", has_255=True)
    mod = load_ref("4BorderClass")
    # raw resized annotation / class maps as the reference's PIL call yields them (input of the label encodings)
    for n in NAMES:
        out["raw/%s/obj" % n] = np.asarray(Image.open(VOC + "SegmentationObject/%s.png" % n).resize(
            (SIZE[0] // 8, SIZE[1] // 8)))
        out["raw/%s/cls" % n] = np.asarray(Image.open(VOC + "SegmentationClass/%s.png" % n).resize(
            (SIZE[0] // 8, SIZE[1] // 8)))
    # A1: the click map for a few sizes / clicks / sigmas
    cases = [((64, 64), (10, 50), 30), ((64, 96), (0, 0), 30), ((96, 64), (95, 63), 20), ((33, 47), (16, 40), 30),
             ((320, 320), (152, 248), 30)]
    for i, (size, where, sigma) in enumerate(cases):
        out["mask_gaussian/%d/args" % i] = np.asarray(list(size) + list(where) + [sigma], np.int64)
        out["mask_gaussian/%d/out" % i] = mod.Data._mask_gaussian(size, list(where), sigma)
    # cfg1: load_image on the reference's own fixture (file name and ndarray input, click given)
    fixture = os.path.join(REF, "input", "7.jpg")
    final, raw, gm = mod.Data.load_image(fixture, where=[40, 25], image_size=(64, 64))[:3]
    out["load_image/file/final"] = final[0]
    out["load_image/file/raw"] = raw
    arr = np.asarray(Image.open(fixture))
    final2 = mod.Data.load_image(arr, where=[12, 60], image_size=(64, 64))[0]
    out["load_image/array/final"] = final2[0]
    np.savez_compressed(os.path.join(HERE, "reference_data.npz"), **out)
    print("wrote reference_data.npz with %d arrays, %d bytes" % (len(out), os.path.getsize(os.path.join(HERE, "reference_data.npz"))))


if __name__ == "__main__":
    sys.exit(main())
