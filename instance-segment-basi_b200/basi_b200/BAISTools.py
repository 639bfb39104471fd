"""Small host utilities with the reference's names (back/2AddClass/BAISTools.py:9-110).

Checkpoints are ``.npz`` files keyed by the TF variable names (conv weights HWIO), so weights can be
exchanged with the oracle or with a TF-side dump; ``restore_if_y`` keeps the resume-if-present contract
of BAISTools.py:95-107.
"""
from __future__ import annotations

import glob
import os
import time

import numpy as np


class Tools(object):

    @staticmethod
    def new_dir(path):
        if not os.path.exists(path):
            os.makedirs(path)
        return path

    @staticmethod
    def print_info(info):
        print("{} {}".format(time.strftime("%H:%M:%S", time.localtime()), info))

    @staticmethod
    def to_txt(data, file_name):
        with open(file_name, "w") as f:
            for one_data in data:
                f.write("{}\n".format(one_data))

    @staticmethod
    def save(engine, checkpoint_path, global_step, max_to_keep=10):
        path = "{}-{}.npz".format(checkpoint_path, global_step)
        np.savez(path, **engine.get_params())
        old = sorted(glob.glob(checkpoint_path + "-*.npz"), key=os.path.getmtime)
        for f in old[:-max_to_keep]:
            os.remove(f)
        return path

    @staticmethod
    def restore_if_y(engine, log_dir, pretrain=None):
        ckpts = sorted(glob.glob(os.path.join(log_dir, "*.npz")), key=os.path.getmtime)
        pretrain = ckpts[-1] if ckpts else pretrain
        if pretrain:
            have = dict(np.load(pretrain))
            cur = engine.get_params()
            cur.update({k: v for k, v in have.items() if k in cur})   # ignore_missing_vars=True
            engine.set_params(cur)
            Tools.print_info("Restored model parameters from {}".format(pretrain))
            return pretrain
        Tools.print_info('No checkpoint file found.')
        return None

    @staticmethod
    def get_shape(tensor):
        return [int(i) for i in list(tensor.shape)[1:]]
