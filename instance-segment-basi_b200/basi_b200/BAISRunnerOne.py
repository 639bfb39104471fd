"""Single-click inference with the reference's ``Runner`` / ``RunnerGUI`` semantics.

back/4BorderClass/BAISRunnerOne.py:13-84 (``Runner.run``: PNG dumps) and
back/4BorderClass/BAISRunnerGUI.py:10-110 (``RunnerGUI``: network kept alive, legacy-bilinear
upsample of the logits to the input size, mask == argmax(sigmoid) == 1).  The interactive matplotlib
loop is out of scope; ``RunnerGUI.click`` is the compute of one loop iteration (the click-to-mask path).
BN runs on batch statistics of the single image (is_training=True in every reference runner).
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np
import torch

from . import _lib
from .BAISData import CategoryNames, Data
from .BAISPSPNet import PSPNet, Placeholder, VARIANTS
from .BAISTools import Tools
from .engine import Engine


class RunnerGUI(object):

    def __init__(self, log_dir, last_pool_size=90, variant="4BorderClass", num_classes=21, num_segment=4,
                 filter_number=32, precision="f16", device=None, seed=0, use_tc=True):
        self.log_dir = log_dir
        self.last_pool_size = last_pool_size
        self.input_size = [self.last_pool_size * 8, self.last_pool_size * 8]
        self.variant = variant
        self.net, self.engine = self.load_net(num_classes, num_segment, filter_number, precision, device, seed,
                                              use_tc)

    def load_net(self, num_classes, num_segment, filter_number, precision, device, seed, use_tc):
        ph = Placeholder((None, self.input_size[0], self.input_size[1], 4))
        net = PSPNet({'data': ph}, is_training=True, num_classes=num_classes, last_pool_size=self.last_pool_size,
                     filter_number=filter_number, num_segment=num_segment, variant=self.variant)
        eng = Engine(net, 1, precision, False, None, device, use_tc)
        eng.init_params(seed)
        if self.log_dir:
            Tools.restore_if_y(eng, self.log_dir)
        eng.enable_click_input(30)
        S = self.input_size
        self.mask_dev = torch.zeros((1, S[0], S[1]), dtype=torch.int32, device=eng.device)
        _, ph_, pw_, nseg = eng.seg_logits.shape
        eng._call(eng.post, "basi_upsample_legacy_argmax", eng.seg_logits.t.data_ptr(), 1, ph_, pw_, nseg, S[0],
                  S[1], self.mask_dev.data_ptr())
        self.pin_img = torch.zeros((1, S[0], S[1], 3), dtype=torch.uint8).pin_memory()
        self.pin_click = torch.zeros((1, 2), dtype=torch.int32).pin_memory()
        return net, eng

    def click(self, image_u8_resized, where):
        """image: uint8 [S,S,3] already resized to input_size; where=[y,x].  Returns (mask uint8 [S,S], class id)."""
        eng = self.engine
        self.pin_img.copy_(torch.from_numpy(np.array(image_u8_resized, dtype=np.uint8)).view(self.pin_img.shape))
        self.pin_click[0, 0], self.pin_click[0, 1] = int(where[0]), int(where[1])
        eng.feed_clicks(self.pin_img, self.pin_click)
        if eng._graph is None:
            eng.capture(train=False)
        eng.replay()
        predict = self.mask_dev.cpu().numpy()[0]
        cls = int(eng.pred_cls.cpu().numpy()[0]) if eng.pred_cls is not None else -1
        segment = np.asarray(np.where(predict == 1, 1, 0), dtype=np.uint8)
        return segment, cls

    @staticmethod
    def click_position(input_size, image_data, point_xy):
        """(x, y) clicked in the displayed image -> [row, col] at the network's input size (BAISRunnerGUI.py:63-64)."""
        return [int(input_size[0] * point_xy[1] / len(image_data)),
                int(input_size[1] * point_xy[0] / len(image_data[0]))]

    @staticmethod
    def mask_to_image_size(segment, image_data):
        """{0,1} uint8 mask at the input size -> the displayed image's size, BAISRunnerGUI.py:84:
        ``Image.fromarray(segment).resize((w, h))`` with PIL's default filter -- no resample argument, like the
        reference, so it is whatever the installed Pillow applies to an 'L' image (bicubic since Pillow 7; NEAREST,
        which this method used before, differs from it on boundary pixels)."""
        from PIL import Image
        return np.asarray(Image.fromarray(segment).resize((len(image_data[0]), len(image_data))))

    @staticmethod
    def blend(image_data, segment, mask_color, opacity):
        """The image the tool displays (BAISRunnerGUI.py:86-97): mask_color blended into the masked pixels."""
        image_mask = np.ndarray(image_data.shape)
        for c in range(3):
            image_mask[:, :, c] = (1 - segment) * image_data[:, :, c] + segment * (
                opacity * mask_color[c] + (1 - opacity) * image_data[:, :, c])
        return image_mask.astype(np.uint8)

    def run_image(self, image_filename_or_data, point_xy):
        """One GUI iteration: original-resolution image + clicked (x, y) -> (mask at original size, class)."""
        from PIL import Image
        image_data = np.array(Image.open(image_filename_or_data)) if isinstance(image_filename_or_data, str) \
            else np.asarray(image_filename_or_data)
        where = self.click_position(self.input_size, image_data, point_xy)
        img = np.asarray(Image.fromarray(image_data.astype(np.uint8)).convert("RGB").resize(
            tuple(self.input_size), Image.BICUBIC), dtype=np.uint8)
        seg, cls = self.click(img, where)
        return self.mask_to_image_size(seg, image_data), cls, where


class Runner(object):

    def __init__(self, log_dir, save_dir, last_pool_size=90, **net_kwargs):
        self.save_dir = Tools.new_dir(save_dir)
        self.log_dir = Tools.new_dir(log_dir)
        self.last_pool_size = last_pool_size
        self.input_size = [self.last_pool_size * 8, self.last_pool_size * 8]
        self.net_kwargs = net_kwargs

    def run(self, result_filename, image_filename, where=None, annotation_filename=None, ann_index=0):
        from PIL import Image
        loaded = Data.load_image(image_filename, where=where, annotation_filename=annotation_filename,
                                 ann_index=ann_index, image_size=self.input_size)
        final_batch_data, data_raw, gaussian_mask = loaded[:3]
        kw = dict(variant="4BorderClass", num_classes=21, num_segment=4, filter_number=32, precision="f16")
        kw.update(self.net_kwargs)
        ph = Placeholder((None, self.input_size[0], self.input_size[1], 4))
        net = PSPNet({'data': ph}, is_training=True, num_classes=kw["num_classes"],
                     last_pool_size=self.last_pool_size, filter_number=kw["filter_number"],
                     num_segment=kw["num_segment"], variant=kw["variant"])
        eng = Engine(net, 1, kw["precision"], False, None, kw.get("device"))
        eng.init_params(kw.get("seed", 0))
        Tools.restore_if_y(eng, self.log_dir)
        eng.feed(np.asarray(final_batch_data, dtype=np.float32))
        eng.forward_device()
        raw_output = eng.seg_logits.t.cpu().numpy()
        sigmoid_output = 1.0 / (1.0 + np.exp(-raw_output))
        predict_output = eng.pred_seg.cpu().numpy()[..., 0] if raw_output.shape[-1] > 1 else \
            (raw_output[..., 0] > 0).astype(np.int32)
        result = dict(raw_output=raw_output, sigmoid_output=sigmoid_output, predict_output=predict_output)
        if eng.cls_logits is not None:
            result["raw_output_classes"] = eng.cls_logits.t.cpu().numpy().reshape(1, -1)
            result["pred_classes"] = eng.pred_cls.cpu().numpy()
            print("{} {} {}".format(result["pred_classes"][0], CategoryNames[result["pred_classes"][0] % 21],
                                    result["raw_output_classes"]))
        nseg = raw_output.shape[-1]
        base = os.path.join(self.save_dir, result_filename)
        Image.fromarray(np.asarray(np.squeeze(data_raw), dtype=np.uint8)).save(base + "data.png")
        if nseg == 1:
            # the one-logit snapshots' dumps (back/2AddClass/BAISRunnerOne.py:53-62): sigmoid as grey levels, the raw
            # logit thresholded at 0.5 (the training-time prediction rule) and the sigmoid thresholded at 0.5
            Image.fromarray(np.asarray(np.squeeze(sigmoid_output[0] * 255), dtype=np.uint8)).save(base + "pred.png")
            Image.fromarray(np.asarray(np.squeeze(np.greater(raw_output[0], 0.5) * 255), dtype=np.uint8)).save(
                base + "pred_raw.png")
            Image.fromarray(np.asarray(np.squeeze(np.greater(sigmoid_output[0], 0.5) * 255), dtype=np.uint8)).save(
                base + "pred_sigmoid.png")
        else:
            Image.fromarray(np.squeeze(np.asarray(predict_output[0] * 255 // nseg, dtype=np.uint8))).save(
                base + "pred.png")
            for c in range(nseg):
                Image.fromarray(np.asarray(sigmoid_output[0, :, :, c] * 255, dtype=np.uint8)).save(
                    base + "pred_%d.png" % c)
        Image.fromarray(np.asarray(np.squeeze(gaussian_mask * 255), dtype=np.uint8)).save(base + "mask.bmp")
        if len(loaded) > 3:
            Image.fromarray(np.asarray(np.squeeze(loaded[4] * 255), dtype=np.uint8)).save(base + "ann.bmp")
        return result
