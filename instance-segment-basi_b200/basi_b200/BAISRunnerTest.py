"""Inference runner of the top-level (current HEAD) path with the reference's call surface.

BAISRunnerTest.py:79-146: ``Inference(input_size, summary_dir, log_dir, model_name).load_model()`` then
``.inference(image_path, image_index, save_path)``: the image is resized to ``input_size`` and scaled to [0, 1]
(BAISData.Data.load_data), the top-level LinkNet runs forward and ``pred_segment = argmax(segments[0])`` -- the
coarsest deep-supervised head -- is written as an 8-bit mask.  TensorBoard summaries (the reference's FileWriter
dumps of every feature channel) are control plane and out of scope; ``summary_dir`` is accepted and ignored.
"""
from __future__ import annotations

import os

import numpy as np
import torch

from .BAISNet import LinkNetTop, Placeholder
from .BAISTools import Tools
from .engine import Engine


class Inference(object):

    def __init__(self, input_size, summary_dir=None, log_dir="./model", model_name="model.ckpt", width=1.0,
                 precision="f16", device=None, seed=0):
        self.log_dir = Tools.new_dir(log_dir)
        self.model_name = model_name
        self.checkpoint_path = os.path.join(self.log_dir, self.model_name)
        self.input_size = input_size
        self.num_classes = 21
        self.image_placeholder = Placeholder((None, self.input_size[0], self.input_size[1], 3))
        self.net = LinkNetTop(self.image_placeholder, False, num_classes=self.num_classes, width=width)
        self.segments, self.features = self.net.build()
        self.engine = Engine(self.net, 1, precision, False, None, device)
        self.engine.init_params(seed)

    def load_model(self):
        return Tools.restore_if_y(self.engine, self.log_dir)

    @staticmethod
    def load_data(image_path, input_size):
        """BAISData.Data.load_data: RGB, resized to input_size, float32 / 255."""
        from PIL import Image
        img = Image.open(image_path) if isinstance(image_path, str) else Image.fromarray(np.asarray(image_path, np.uint8))
        return np.asarray(img.convert("RGB").resize(tuple(input_size)), dtype=np.float32) / 255

    def predict(self, im_data):
        """im_data: float32 [S,S,3] in [0,1] -> uint8 mask of the coarsest head (argmax over its 2 logits)."""
        eng = self.engine
        eng.feed(np.expand_dims(np.asarray(im_data, dtype=np.float32), 0))
        eng.forward_device()
        torch.cuda.synchronize(eng.device)
        logits = eng.att_logits[0].t.float().cpu().numpy()[0]
        return np.argmax(logits, axis=-1).astype(np.uint8)

    def inference(self, image_path, image_index=0, save_path=None):
        from PIL import Image
        pred = self.predict(self.load_data(image_path, self.input_size))
        s_image = Image.fromarray(np.asarray(pred * 255, dtype=np.uint8))
        if save_path is not None:
            Tools.new_dir(save_path)
            s_image.convert("L").save("{}/{}.bmp".format(save_path, os.path.splitext(os.path.basename(image_path))[0]))
        return pred
