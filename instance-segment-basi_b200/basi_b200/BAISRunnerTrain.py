"""Training runner with the reference's ``Train`` call surface.

back/2AddClass/BAISRunnerTrain.py:13-188 (segment + attention-class, pos_weight BCE, 0.1/0.2 class term),
back/1NoClass (segment only), back/4BorderClass and back/5COCO (softmax segment heads): same constructor
kwargs and ``train(save_pred_freq, begin_step)``; ``variant`` picks the snapshot.  One ``sess.run`` of the
reference == one ``Engine`` step here (forward, fused losses, backward, SGD on the B200).
"""
from __future__ import annotations

import os
import time

import numpy as np

from .BAISData import Data, DataAttention, DataTop, SyntheticData
from .BAISPSPNet import PSPNet, Placeholder, VARIANTS
from .BAISTools import Tools
from .engine import Engine

# per-snapshot training constants (file:line in the class docstring above)
SNAPSHOT = {
    "1NoClass": dict(num_segment=1, kind="bce", pos_weight=3.0, class_weight=0.0, lr=1e-2, num_steps=400001),
    "2AddClass": dict(num_segment=1, kind="bce", pos_weight=3.0, class_weight=0.2, lr=5e-3, num_steps=500001),
    "3ThreeClass": dict(num_segment=3, kind="softmax", pos_weight=1.0, class_weight=0.1, lr=5e-3, num_steps=500001),
    "4BorderClass": dict(num_segment=4, kind="softmax", pos_weight=1.0, class_weight=0.1, lr=5e-3, num_steps=500001),
    "5COCO": dict(num_segment=3, kind="softmax", pos_weight=1.0, class_weight=0.2, lr=5e-3, num_steps=1000001),
    # variant B (back/90AttentionSingle2/BAISRunnerTrain.py:30-52,116-156): vgg_16 trunk + attention cascade,
    # loss = mean 2-channel weighted CE of the four attention maps + class CE
    "90AttentionSingle2": dict(num_segment=1, kind="linknet_b", pos_weight=3.0, class_weight=1.0, lr=5e-4,
                               num_steps=500001),      # (:26 learning_rate = 5e-4 -- a tenth of the PSPNet snapshots')
    # cascaded attention re-decoding (back/8AttentionU/BAISRunnerTrain.py:28-38,161-193): 2AddClass trunk + four
    # pyramid decoders; loss on the sigmoid outputs (2 x softmax CE, 2 x doubled weighted BCE) + 0.1 * four class CEs
    "8AttentionU": dict(num_segment=4, kind="cascade", pos_weight=3.0, class_weight=0.1, lr=5e-3, num_steps=500001),
}

# The current-HEAD training script (BAISRunnerTrain.py:28-50,97-115: BAISNet.LinkNet on 3-channel images, full-resolution
# {0,1} labels, cal_loss = mean of five 2-channel weighted CEs with pos_weight 1) runs on a DIFFERENT schedule: power 0.8
# over 100 001 steps.  Its network / loss / optimizer subset are `BAISNet.LinkNetTop` + `Engine(kind="linknet_b",
# pos_weight=1.0)` + `Engine.set_trainable("segment_side")`; `TrainTop` below wraps them with that script's call surface
# (`Train` itself is the PSPNet-lineage / variant-B / cascade runner: click input and class labels).
HEAD_SCHEDULE = dict(base_lr=5e-3, num_steps=100001, power=0.8)


def poly_learning_rate(base_lr, step, num_steps, power=0.9):
    """lr = base * (1 - step/num_steps)^0.9 in float32, step fed as float32 (BAISRunnerTrain.py:115-116)."""
    s = np.float32(step) / np.float32(num_steps)
    return float(np.float32(base_lr) * np.power(np.float32(1) - s, np.float32(power)))


class Train(object):

    def __init__(self, batch_size, last_pool_size, input_size, log_dir, data_root_path=None, train_list=None,
                 data_path=None, annotation_path=None, class_path=None, model_name="model.ckpt", is_test=False,
                 variant="2AddClass", num_classes=21, precision="f16", filter_number=32, pos_weight=None,
                 class_weight=None, learning_rate=None, num_steps=None, seed=0, device=None, dp=None,
                 use_cuda_graph=True, use_tc=True):
        snap = SNAPSHOT[variant]
        self.log_dir = Tools.new_dir(log_dir)
        self.model_name = model_name
        self.checkpoint_path = os.path.join(self.log_dir, self.model_name)
        self.input_size = input_size
        self.batch_size = batch_size
        self.num_classes = num_classes
        self.num_segment = snap["num_segment"]
        self.ratio = 8
        self.last_pool_size = last_pool_size
        self.filter_number = filter_number
        self.learning_rate = snap["lr"] if learning_rate is None else learning_rate
        self.num_steps = snap["num_steps"] if num_steps is None else num_steps
        self.variant = variant
        self.print_step = 1 if is_test else 25
        self.dp = dp
        self.use_cuda_graph = use_cuda_graph
        self.synthetic = data_root_path is None
        if self.synthetic:
            self.data_reader = SyntheticData(batch_size, tuple(input_size), self.ratio, num_classes,
                                             self.num_segment, sigma=20 if variant == "5COCO" else 30,
                                             seed=seed + (dp.rank if dp else 0))
        elif variant == "90AttentionSingle2":
            # variant B reads full-resolution attention labels and feeds image and click map separately
            # (back/90AttentionSingle2/BAISRunnerTrain.py:31-34,177-180)
            self.data_reader = DataAttention(data_root_path=data_root_path, data_list=train_list, data_path=data_path,
                                             annotation_path=annotation_path, class_path=class_path,
                                             batch_size=batch_size, image_size=input_size, is_test=is_test,
                                             rank=dp.rank if dp else 0, world=dp.world if dp else 1, seed=seed)
        else:
            self.data_reader = Data(data_root_path=data_root_path, data_list=train_list, data_path=data_path,
                                    annotation_path=annotation_path, class_path=class_path,
                                    batch_size=batch_size, image_size=input_size, is_test=is_test,
                                    rank=dp.rank if dp else 0, world=dp.world if dp else 1, seed=seed)
        self.loss_cfg = dict(kind=snap["kind"],
                             pos_weight=snap["pos_weight"] if pos_weight is None else pos_weight,
                             class_weight=snap["class_weight"] if class_weight is None else class_weight)
        self.net, self.engine = self.build_net(precision, device, use_tc)
        self.engine.init_params(seed)
        if dp is not None:
            self.engine.broadcast_params(dp)       # every replica starts from rank 0's weights

    def build_net(self, precision="f16", device=None, use_tc=True):
        if self.variant == "90AttentionSingle2":
            from .BAISNet import LinkNet
            net = LinkNet(Placeholder((None, self.input_size[0], self.input_size[1], 3)),
                          Placeholder((None, self.input_size[0], self.input_size[1], 1), name="mask"),
                          is_training=True, num_classes=self.num_classes, width=self.filter_number / 64.0)
            if not self.synthetic:
                # the reference's reader delivers FULL-resolution labels (label_seg_placeholder [S, S, 1],
                # back/90AttentionSingle2/BAISRunnerTrain.py:40) that cal_loss nearest-resizes to every attention map;
                # the synthetic reader's labels are S/8 maps.  (First GPU run of this real-data path -- session 3, last
                # seconds of the budget -- stopped here on the S/8 label buffer; with label_stride = 1 it takes the label
                # path the top-level LinkNet's tests exercise.  Not re-run on the GPU.)
                net.label_stride = 1
            engine = Engine(net, self.batch_size, precision, True, self.loss_cfg, device, use_tc)
            if self.synthetic:
                engine.enable_click_input(self.data_reader.sigma)
            self.raw_output_segment = net.attentions[-1]
            self.raw_output_classes = net.classes[0]
            return net, engine
        if self.variant == "8AttentionU":
            from .BAISNet import BAISNet
            net = BAISNet(Placeholder((None, self.input_size[0], self.input_size[1], 4)), is_training=True,
                          num_classes=self.num_classes, num_segment=self.num_segment, segment_attention=1,
                          last_pool_size=self.last_pool_size, filter_number=self.filter_number,
                          attention_module_num=2)
            engine = Engine(net, self.batch_size, precision, True, self.loss_cfg, device, use_tc)
            if self.synthetic:
                engine.enable_click_input(self.data_reader.sigma)
            self.segments, self.attentions, self.classes = net.build()
            self.raw_output_segment = self.segments[0]           # final_segment_logit / final_class_logit (:63-64)
            self.raw_output_classes = self.classes[0]
            return net, engine
        image_placeholder = Placeholder((None, self.input_size[0], self.input_size[1], 4))
        net = PSPNet({'data': image_placeholder}, is_training=True, num_classes=self.num_classes,
                     num_segment=self.num_segment, last_pool_size=self.last_pool_size,
                     filter_number=self.filter_number, variant=self.variant)
        engine = Engine(net, self.batch_size, precision, True, self.loss_cfg, device, use_tc)
        if self.synthetic:
            engine.enable_click_input(self.data_reader.sigma)
        self.raw_output_segment = net.layers[VARIANTS[self.variant]["seg"]]
        fc = VARIANTS[self.variant]["fc"]
        self.raw_output_classes = net.layers[fc] if fc else None
        return net, engine

    def _attention_labels(self, label_seg):
        """8AttentionU: ann_attention = (ann == 1) next to the 4-class labels (back/8AttentionU/BAISData.py:67)."""
        if self.variant != "8AttentionU":
            return None
        return (np.asarray(label_seg) == 1).astype(np.float32)

    def run_step(self, step, batch=None, fetch=True, prefetch=None):
        """One reference ``sess.run([... train_op ...], feed_dict)``; returns the fetched values.

        `prefetch` (synthetic / click-input batches): the batch of the NEXT call -- its host-to-device copies are
        started on a copy stream as soon as this step is enqueued, so they overlap it; the next call must then pass
        that same object as `batch`."""
        eng = self.engine
        lr = poly_learning_rate(self.learning_rate, step, self.num_steps)
        staged = getattr(self, "_staged_batch", None)
        self._staged_batch = None
        if self.synthetic:
            images, clicks, label_seg, label_cls = batch if batch is not None else self.data_reader.next_batch()
            if staged is not None and batch is staged:
                eng.commit_staged()
                eng.set_lr(lr)
            else:
                eng.feed_clicks(images, clicks)
                eng.feed(None, label_seg, label_cls, lr, label_att=self._attention_labels(label_seg))
        elif self.variant == "90AttentionSingle2":
            data, mask, ann, cls = batch if batch is not None else self.data_reader.next_batch_train()
            label_cls, label_seg = cls, np.asarray(ann)
            eng.feed(np.asarray(data, dtype=np.float32), label_seg, np.asarray(cls, dtype=np.int32), lr,
                     mask=np.asarray(mask, dtype=np.float32))
        else:
            data, ann, cls, _, _ = batch if batch is not None else self.data_reader.next_batch_train()
            label_cls, label_seg = cls, np.asarray(ann)
            eng.feed(np.asarray(data, dtype=np.float32), label_seg, np.asarray(cls, dtype=np.int32), lr,
                     label_att=self._attention_labels(label_seg))
        if self.use_cuda_graph:
            if eng._graph is None:
                eng.capture(train=True, sync_grads=self.dp)
            eng.replay()
        elif self.dp is not None:
            eng.step_device(sync_grads=self.dp)
        else:
            eng.step_device()
        if prefetch is not None and self.synthetic:
            p_img, p_clicks, p_seg, p_cls = prefetch
            eng.stage_batch(p_img, p_clicks, p_seg, p_cls, self._attention_labels(p_seg))
            self._staged_batch = prefetch
        if not fetch:
            return None
        att = ([("raw_output_attentions_%d" % i, a.t) for i, a in enumerate(eng.att_logits)]
               if self.variant == "90AttentionSingle2" else [])
        got = eng.fetch_step(att)          # one synchronisation for the whole fetch list
        loss, loss_seg, loss_cls = eng.losses_from(got["loss_acc"])
        raw, pred_seg = got["raw_output_segment"], got["pred_segment"]
        out = dict(loss=loss, loss_segment=loss_seg, loss_classes=loss_cls, learning_rate=lr,
                   raw_output_segment=raw, pred_segment=pred_seg)
        if self.variant == "90AttentionSingle2":
            out["raw_output_attentions"] = [got[k] for k, _ in att]
            lab = None
        else:
            lab = np.asarray(label_seg).reshape(pred_seg.shape)
        if lab is not None:
            out.update(self.segment_accuracies(self.num_segment, raw, pred_seg, lab))
        if eng.cls_logits is not None:
            out["raw_output_classes"] = got["raw_output_classes"].reshape(self.batch_size, -1)
            out["pred_classes"] = got["pred_classes"]
            out["accuracy_classes"] = self.class_accuracy(out["pred_classes"], label_cls)
        return out

    @staticmethod
    def segment_accuracies(num_segment, raw_output_segment, pred_segment, label_segment):
        """The per-batch statistics build_net calls accuracies: one-logit snapshots report the FRACTION of positive
        predictions / positive labels (back/2AddClass/BAISRunnerTrain.py:96-97: accuracy_0 = mean(logit > 0.5),
        accuracy_1 = mean(label > 0.5)), the softmax snapshots tcm.accuracy(pred_segment, label)
        (back/4BorderClass/BAISRunnerTrain.py:102)."""
        if num_segment == 1:
            return dict(accuracy_0=float(np.mean(np.asarray(raw_output_segment).reshape(-1) > 0.5)),
                        accuracy_1=float(np.mean(np.asarray(label_segment).reshape(-1) > 0.5)))
        lab = np.asarray(label_segment).reshape(np.asarray(pred_segment).shape)
        return dict(accuracy_segment=float(np.mean(np.asarray(pred_segment) == lab)))

    @staticmethod
    def class_accuracy(pred_classes, label_classes):
        """tcm.accuracy(pred_classes, label_classes_placeholder) (back/2AddClass/BAISRunnerTrain.py:98)."""
        return float(np.mean(np.asarray(pred_classes) == np.asarray(label_classes)))

    def train(self, save_pred_freq, begin_step=0, max_steps=None):
        Tools.restore_if_y(self.engine, self.log_dir)
        if self.dp is not None:
            self.engine.broadcast_params(self.dp)
        is_writer = self.dp is None or self.dp.rank == 0     # one checkpoint writer
        end = self.num_steps if max_steps is None else min(self.num_steps, begin_step + max_steps)
        r = None
        # synthetic / click-input batches are drawn one step ahead and prefetched (same draw order as without)
        nxt = self.data_reader.next_batch() if (self.synthetic and begin_step < end) else None
        for step in range(begin_step, end):
            start_time = time.time()
            cur, nxt = nxt, (self.data_reader.next_batch() if (self.synthetic and step + 1 < end) else None)
            r = self.run_step(step, cur, prefetch=nxt)
            if step % save_pred_freq == 0 and is_writer:
                Tools.save(self.engine, self.checkpoint_path, step)
                Tools.print_info('The checkpoint has been created.')
            duration = time.time() - start_time
            if step % self.print_step == 0:
                Tools.print_info('step {:d} loss={:.3f} seg={:.3f} class={:.3f} lr={:.6f} ({:.3f} s/step)'.format(
                    step, r["loss"], r["loss_segment"], r["loss_classes"], r["learning_rate"], duration))
        return r


class TrainTop(object):
    """The current-HEAD training script with its call surface (BAISRunnerTrain.py:10-183:
    ``Train(batch_size, input_size, log_dir, data_root_path, train_list, data_path, annotation_path, class_path,
    model_name, pretrain, is_test).train(save_pred_freq, begin_step)``): `DataTop` batches of 3-channel images and
    full-resolution foreground maps -> ``BAISNet.LinkNetTop`` -> ``cal_loss`` (mean of the five 2-channel weighted CEs,
    pos_weight 1) -> plain SGD on the HEAD schedule (power 0.8 over 100 001 steps); ``segment_side_only=True`` is the
    script's ``train_segment_side_op`` (``minimize(loss, var_list=[... 'segment_side' ...])``).

    Host glue only: every device call below (``Engine.feed`` of image + full-resolution labels, ``step_device`` /
    graph replay, ``losses``) is the one the top-level LinkNet's GPU tests make.  The loop itself was added after the GPU
    budget of round 2 was spent: plan construction, feeds and the schedule are tested on the CPU (dry run), the loop has
    not been run on a GPU."""

    def __init__(self, batch_size, input_size, log_dir, data_root_path=None, train_list=None, data_path=None,
                 annotation_path=None, class_path=None, model_name="model.ckpt", pretrain=None, is_test=False,
                 precision="f16", width=1.0, seed=0, device=None, use_cuda_graph=True, dry_run=False):
        from .BAISNet import LinkNetTop
        self.log_dir = Tools.new_dir(log_dir)
        self.model_name = model_name
        self.checkpoint_path = os.path.join(self.log_dir, self.model_name)
        self.pretrain = pretrain
        self.input_size = input_size
        self.batch_size = batch_size
        self.num_classes = 21
        self.data_reader = DataTop(data_root_path=data_root_path, data_list=train_list, data_path=data_path,
                                   annotation_path=annotation_path, class_path=class_path, batch_size=batch_size,
                                   image_size=input_size, is_test=is_test, seed=seed)
        self.learning_rate, self.num_steps = HEAD_SCHEDULE["base_lr"], HEAD_SCHEDULE["num_steps"]
        self.cal_step = self.data_reader.number_patch
        self.print_step = max(1, self.cal_step // 10)
        self.use_cuda_graph = use_cuda_graph
        self.net = LinkNetTop(Placeholder((None, input_size[0], input_size[1], 3)), True,
                              num_classes=self.num_classes, width=width)
        self.segments, self.features = self.net.build()
        self.engine = Engine(self.net, batch_size, precision, True, dict(kind="linknet_b", pos_weight=1.0), device,
                             dry_run=dry_run)
        self.engine.init_params(seed)
        self._subset = None

    def feed(self, step, batch=None):
        """Host -> static device buffers for one step; returns the learning rate of the step."""
        lr = poly_learning_rate(step=step, **HEAD_SCHEDULE)
        data, ann = batch if batch is not None else self.data_reader.next_batch_train()
        self.engine.feed(np.asarray(data, dtype=np.float32), np.asarray(ann, dtype=np.float32), None, lr)
        return lr

    def run_step(self, step, batch=None, segment_side_only=False):
        """One ``sess.run([train_op | train_segment_side_op, loss, loss_segment_all, learning_rate])``."""
        eng = self.engine
        subset = "segment_side" if segment_side_only else None
        if subset != self._subset:
            eng.set_trainable(subset)
            self._subset = subset
        lr = self.feed(step, batch)
        if self.use_cuda_graph:
            if eng._graph is None:
                eng.capture(train=True)
            eng.replay()
        else:
            eng.step_device()
        _, loss_segment_all, _ = eng.losses()
        return dict(loss=loss_segment_all, loss_segment_all=loss_segment_all, learning_rate=lr)

    def train(self, save_pred_freq, begin_step=0, max_steps=None):
        Tools.restore_if_y(self.engine, self.log_dir, pretrain=self.pretrain)
        end = self.num_steps if max_steps is None else min(self.num_steps, begin_step + max_steps)
        r = None
        for step in range(begin_step, end):
            start_time = time.time()
            r = self.run_step(step)
            if step % save_pred_freq == 0:
                Tools.save(self.engine, self.checkpoint_path, step)
                Tools.print_info('The checkpoint has been created.')
            if step % self.print_step == 0:
                Tools.print_info('step {:d} loss={:.3f} lr={:.6f} ({:.3f} s/step)'.format(
                    step, r["loss"], r["learning_rate"], time.time() - start_time))
        return r
