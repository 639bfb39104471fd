"""Data side of the hot path: click sampling, Gaussian click map, batch assembly.

Call surface of back/2AddClass/BAISData.py (``Data`` :16-252): ``next_batch_train()`` returns the
same 5-tuple, ``Data._mask_gaussian`` / ``Data.load_image`` keep their signatures.  The per-step
numpy work of the reference (B float64 exp maps + np.concatenate on one host thread, :70-80) moves
to the GPU: ``click_lut`` builds the d^2-indexed table once and ``basi_clickmap_pack`` gathers it,
which is bit-exact with the numpy expression by construction (d^2 is an exact integer).

``SyntheticData`` produces VOC/COCO-shaped batches without any dataset on disk (benchmarks, tests).
"""
from __future__ import annotations

import numpy as np

CategoryNames = ['background',
                 'aeroplane', 'bicycle', 'bird', 'boat', 'bottle',
                 'bus', 'car', 'cat', 'chair', 'cow',
                 'diningtable', 'dog', 'horse', 'motorbike', 'person',
                 'pottedplant', 'sheep', 'sofa', 'train', 'tvmonitor']


def click_lut(image_size, sigma=30):
    """float32 table t[d2] == float32(exp(-4 ln2 d2 / sigma^2)) for every reachable squared distance."""
    h, w = int(image_size[0]), int(image_size[1])
    d2 = np.arange(0, (h - 1) ** 2 + (w - 1) ** 2 + 1, 1, float)
    return np.exp(-4 * np.log(2) * d2 / sigma ** 2).astype(np.float32)


class Data(object):
    """VOC reader with the reference's interface (PIL + numpy on the host)."""

    def __init__(self, data_list="ImageSets/Segmentation/train.txt", data_path="JPEGImages/",
                 data_root_path="./VOC2012/", annotation_path="SegmentationObject/",
                 class_path="SegmentationClass/", batch_size=4, image_size=(720, 720), ratio=8, is_test=False,
                 sigma=30, rank=0, world=1, seed=None):
        # data parallelism: every rank reads the annotations rank::world of one common (per-epoch seeded) order, so
        # the ranks see disjoint batches; rank=0, world=1 is the reference's single-process behaviour
        self.rank, self.world = int(rank), max(1, int(world))
        self._epoch, self._seed = 0, seed
        self.batch_size = batch_size
        self.image_size = image_size
        self.ratio = ratio
        self.sigma = sigma
        self._data_list, self._annotation_list, self._class_list = self._read_list(
            data_root_path, data_list, data_path, annotation_path, class_path)
        if is_test:
            self._data_list, self._annotation_list = self._data_list[0: 12], self._annotation_list[0: 12]
            self._class_list = self._class_list[0: 12]
        cut = getattr(self, "_is_test_cut", None)          # (the full-resolution readers keep the first 200 entries)
        if cut:
            self._data_list = self._data_list[0: cut]
        self._annotations = self._read_annotation(self._annotation_list, self._class_list, self.image_size,
                                                  self.ratio)
        self._images_data = self._read_image(self._data_list, self.image_size)
        self._order = list(range(0, len(self._annotations)))
        self._random_index = self._order[self.rank::self.world]
        self.number_patch = len(self._random_index) // self.batch_size
        self._now = 0

    def _reshuffle(self):
        self._epoch += 1
        if self.world == 1 and self._seed is None:
            np.random.shuffle(self._random_index)            # the reference's in-place shuffle (BAISData.py:52-54)
            return
        order = list(self._order)
        np.random.RandomState((0 if self._seed is None else self._seed) + self._epoch).shuffle(order)
        self._random_index = order[self.rank::self.world]    # same permutation on every rank, disjoint slices

    def _pick(self, train):
        if train and self._now >= self.number_patch:
            self._reshuffle()
            self._now = 0
        idx = self._random_index[self._now * self.batch_size: (self._now + 1) * self.batch_size]
        return [self._annotations[i] for i in idx]

    def next_batch_train(self):
        batch_ann = self._pick(True)
        for ann in batch_ann:
            where = np.argwhere(ann[-1] == 1)
            where = where[np.random.randint(0, len(where))]
            ann[1] = [where[0] * self.ratio, where[1] * self.ratio]
        batch_mask = [self._mask_gaussian(self.image_size, ann[1], self.sigma) for ann in batch_ann]
        batch_data = [self._images_data[ann[0]] for ann in batch_ann]
        final_batch_data = [np.concatenate((d, np.expand_dims(m, 2)), 2) for d, m in zip(batch_data, batch_mask)]
        final_batch_ann = [np.expand_dims(ann[-1], 2) for ann in batch_ann]
        final_batch_class = [ann[2] for ann in batch_ann]
        self._now += 1
        return final_batch_data, final_batch_ann, final_batch_class, batch_data, batch_mask

    @staticmethod
    def _read_annotation(annotation_list, class_list, image_size, ratio):
        from PIL import Image
        out = []
        size = (image_size[0] // ratio, image_size[1] // ratio)
        for ann_index, ann_name in enumerate(annotation_list):
            class_data = np.asarray(Image.open(class_list[ann_index]).resize(size, Image.NEAREST))
            ann_data = np.asarray(Image.open(ann_name).resize(size, Image.NEAREST))
            for num in [i for i in range(1, 255) if np.any(ann_data == i)]:
                ys, xs = np.where(ann_data == num)
                cls = class_data[ys[0]][xs[0]]
                cls = 0 if cls >= len(CategoryNames) else int(cls)
                out.append([ann_index, num, cls, np.where(ann_data == num, 1, 0)])
        return out

    @staticmethod
    def _read_image(data_list, image_size):
        from PIL import Image
        out = []
        for name in data_list:
            d = np.asarray(Image.open(name).convert("RGB").resize(image_size, Image.BICUBIC), dtype=np.float32)
            d /= 255
            out.append(d)
        return out

    @staticmethod
    def _read_list(data_root_path, data_list, data_path, annotation_path, class_path):
        with open(data_root_path + data_list, "r") as f:
            names = [line.strip() for line in f.readlines()]
        return ([data_root_path + data_path + n + ".jpg" for n in names],
                [data_root_path + annotation_path + n + ".png" for n in names],
                [data_root_path + class_path + n + ".png" for n in names])

    @staticmethod
    def _mask_gaussian(image_size, where, sigma=30):
        """Host version (float64 -> float32), identical to the reference expression."""
        x = np.arange(0, image_size[1], 1, float)
        y = np.arange(0, image_size[0], 1, float)[:, np.newaxis]
        x0, y0 = where[1], where[0]
        return np.exp(-4 * np.log(2) * ((x - x0) ** 2 + (y - y0) ** 2) / sigma ** 2).astype(np.float32)

    @staticmethod
    def load_image(image_filename, where=None, annotation_filename=None, ann_index=0, image_size=(720, 720)):
        """Returns (final_batch_data, data_raw, gaussian_mask[, ann_data, ann_mask]) like the reference."""
        from PIL import Image
        if isinstance(image_filename, str):
            img = Image.open(image_filename)
        else:
            img = Image.fromarray(np.asarray(image_filename, dtype=np.uint8))
        data_raw = np.asarray(img.convert("RGB").resize(tuple(image_size), Image.BICUBIC), dtype=np.float32)
        data_data = data_raw / 255
        extra = ()
        if annotation_filename is not None:
            ann_data = np.asarray(Image.open(annotation_filename).resize(tuple(image_size), Image.NEAREST))
            nums = [i for i in range(1, 255) if np.any(ann_data == i)]
            ann_mask = [np.where(ann_data == i, 1, 0) for i in nums]
            where = np.argwhere(ann_mask[ann_index] == 1)
            where = where[np.random.randint(0, len(where))]
            extra = (ann_data, ann_mask[ann_index])
        if where is None:
            raise Exception("where can not none")
        gaussian_mask = Data._mask_gaussian(image_size, where)
        final_batch_data = [np.concatenate((data_data, np.expand_dims(gaussian_mask, 2)), 2)]
        return (final_batch_data, data_raw, gaussian_mask) + extra


class DataAttention(Data):
    """Reader of variant B (back/90AttentionSingle2/BAISData.py:17-125): instance masks at the FULL input resolution with
    the 255 border ring counted as background, the click drawn from the mask at full resolution (no ratio), and
    ``next_batch_train() -> (batch_data [S,S,3] f32, batch_mask [S,S,1] f32 click map, batch_attention [S,S,1] int32
    {0,1}, batch_class)`` -- image and click map are separate feeds of ``LinkNet(input_data, input_mask, ...)``."""

    def __init__(self, data_list="ImageSets/Segmentation/trainval.txt", data_path="JPEGImages/",
                 data_root_path="./VOC2012/", annotation_path="SegmentationObject/", class_path="SegmentationClass/",
                 batch_size=4, image_size=(720, 720), is_test=False, sigma=30, rank=0, world=1, seed=None):
        self._is_test_cut = 200 if is_test else None       # (:32-35: the first 200 list entries)
        Data.__init__(self, data_list, data_path, data_root_path, annotation_path, class_path, batch_size, image_size,
                      1, False, sigma, rank, world, seed)

    def _read_annotation(self, annotation_list, class_list, image_size, ratio):
        from PIL import Image
        if self._is_test_cut:
            annotation_list, class_list = annotation_list[:self._is_test_cut], class_list[:self._is_test_cut]
        out = []
        size = (image_size[0], image_size[1])
        for ann_index, ann_name in enumerate(annotation_list):
            class_data = np.asarray(Image.open(class_list[ann_index]).resize(size, Image.NEAREST))
            ann_data = np.asarray(Image.open(ann_name).resize(size, Image.NEAREST))
            ann_data = np.where(ann_data == 255, 0, ann_data)                     # the border ring is background (:93-94)
            for num in [i for i in range(1, 255) if np.any(ann_data == i)]:
                ys, xs = np.where(ann_data == num)
                cls = class_data[ys[0]][xs[0]]
                cls = 0 if cls >= len(CategoryNames) else cls
                out.append([ann_index, num, cls, np.where(ann_data == num, 255, ann_data) // 255])    # (:104)
        return out

    def next_batch_train(self):
        batch_ann = self._pick(True)
        batch_ann_attention = [np.asarray(ann[-1] == 1, dtype=np.int32) for ann in batch_ann]
        for ann in batch_ann:
            where = np.argwhere(ann[-1] == 1)
            where = where[np.random.randint(0, len(where))]
            ann[1] = [where[0], where[1]]
        batch_data = [self._images_data[ann[0]] for ann in batch_ann]
        batch_mask = [np.expand_dims(self._mask_gaussian(self.image_size, ann[1], self.sigma), axis=-1)
                      for ann in batch_ann]
        batch_attention = [np.expand_dims(a, 2) for a in batch_ann_attention]
        batch_class = [ann[2] for ann in batch_ann]
        self._now += 1
        return batch_data, batch_mask, batch_attention, batch_class


class DataTop(Data):
    """Reader of the current-HEAD script (BAISData.py:17-128): ONE foreground map per image at the full input resolution
    -- every labelled pixel including the 255 border ring is 1 (:86-90) -- no clicks, no classes;
    ``next_batch_train() -> (batch_data [S,S,3] f32, batch_ann [S,S,1] {0,1})``."""

    def __init__(self, data_list="ImageSets/Segmentation/trainval.txt", data_path="JPEGImages/",
                 data_root_path="./VOC2012/", annotation_path="SegmentationObject/", class_path="SegmentationClass/",
                 batch_size=4, image_size=(720, 720), is_test=False, rank=0, world=1, seed=None):
        self._is_test_cut = 200 if is_test else None
        Data.__init__(self, data_list, data_path, data_root_path, annotation_path, class_path, batch_size, image_size,
                      1, False, 30, rank, world, seed)

    def _read_annotation(self, annotation_list, class_list, image_size, ratio):
        from PIL import Image
        if self._is_test_cut:
            annotation_list = annotation_list[:self._is_test_cut]
        out = []
        for ann_name in annotation_list:
            ann_data = np.asarray(Image.open(ann_name).resize((image_size[0], image_size[1]), Image.NEAREST))
            ann_data = np.where(ann_data == 255, 1, ann_data)
            out.append(np.where(ann_data > 0, 1, 0))
        return out

    def _pick_images(self):
        if self._now >= self.number_patch:
            self._reshuffle()
            self._now = 0
        return self._random_index[self._now * self.batch_size: (self._now + 1) * self.batch_size]

    def next_batch_train(self):
        idx = self._pick_images()
        batch_data = [self._images_data[i] for i in idx]
        batch_ann = [np.expand_dims(self._annotations[i], axis=-1) for i in idx]
        self._now += 1
        return batch_data, batch_ann


class SyntheticData(object):
    """Seeded VOC/COCO-shaped batches: uint8 images, one filled ellipse or box per
    sample as the instance at P x P, optional border ring / distractor for the 4-class encoding, a click
    drawn from the instance pixels times the ratio, class ids U{1..num_classes-1}."""

    def __init__(self, batch_size, image_size=(320, 320), ratio=8, num_classes=21, num_segment=1, sigma=30, seed=0,
                 coco=False):
        self.batch_size, self.image_size, self.ratio = batch_size, tuple(image_size), ratio
        self.num_classes, self.num_segment, self.sigma, self.coco = num_classes, num_segment, sigma, coco
        self.rng = np.random.RandomState(seed)

    def _label(self):
        ph, pw = self.image_size[0] // self.ratio, self.image_size[1] // self.ratio
        yy, xx = np.mgrid[0:ph, 0:pw]
        rng = self.rng
        cy, cx = rng.uniform(0.3, 0.7) * ph, rng.uniform(0.3, 0.7) * pw
        ry, rx = rng.uniform(0.15, 0.35) * ph, rng.uniform(0.15, 0.35) * pw
        if rng.rand() < 0.5:
            inst = ((yy - cy) / ry) ** 2 + ((xx - cx) / rx) ** 2 <= 1.0
        else:
            inst = (np.abs(yy - cy) <= ry) & (np.abs(xx - cx) <= rx)
        inst[int(cy), int(cx)] = True
        if self.num_segment == 1:
            return inst.astype(np.float32), inst
        lab = np.zeros((ph, pw), dtype=np.int32)
        if self.num_segment == 4:      # 0 other, 1 attention, 2 border, 3 background
            lab[:] = 3
            other = ((yy - 0.15 * ph) ** 2 + (xx - 0.85 * pw) ** 2) <= (0.1 * ph) ** 2
            lab[other] = 0
            grow = np.zeros_like(inst)
            grow[1:, :] |= inst[:-1, :]; grow[:-1, :] |= inst[1:, :]
            grow[:, 1:] |= inst[:, :-1]; grow[:, :-1] |= inst[:, 1:]
            lab[grow & ~inst] = 2
            lab[inst] = 1
        else:                          # 5COCO: 0 background, 1 other, 2 attention
            other = ((yy - 0.15 * ph) ** 2 + (xx - 0.85 * pw) ** 2) <= (0.1 * ph) ** 2
            lab[other] = 1
            lab[inst] = 2
        return lab, inst

    def next_batch(self):
        """(images uint8 [B,H,W,3], clicks int32 [B,2], label_seg [B,P,P,1], label_cls int32 [B])"""
        B, (H, W) = self.batch_size, self.image_size
        images = self.rng.randint(0, 256, size=(B, H, W, 3), dtype=np.uint8)
        labs, clicks = [], []
        for _ in range(B):
            lab, inst = self._label()
            where = np.argwhere(inst)
            w = where[self.rng.randint(0, len(where))]
            clicks.append([w[0] * self.ratio, w[1] * self.ratio])
            labs.append(lab[..., None])
        label_seg = np.stack(labs)
        label_cls = self.rng.randint(1, self.num_classes, size=(B,)).astype(np.int32)
        return images, np.asarray(clicks, dtype=np.int32), label_seg, label_cls

    def next_batch_train(self):
        """Reference-shaped 5-tuple (host click maps, float images): the slow path, kept for drop-in use."""
        images, clicks, label_seg, label_cls = self.next_batch()
        batch_data = [im.astype(np.float32) / 255 for im in images]
        batch_mask = [Data._mask_gaussian(self.image_size, c, self.sigma) for c in clicks]
        final = [np.concatenate((d, np.expand_dims(m, 2)), 2) for d, m in zip(batch_data, batch_mask)]
        return final, list(label_seg), list(label_cls), batch_data, batch_mask
