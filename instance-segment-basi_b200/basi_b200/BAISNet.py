"""Variant B: slim ``vgg_16`` trunk + click-gated attention cascade, with the reference's ``LinkNet`` call surface.

back/90AttentionSingle2/BAISNet.py:653-810: ``LinkNet(input_data, input_mask, is_training, num_classes).build()``
returns ``(segments, attentions, classes)``.  Here ``input_data`` / ``input_mask`` are ``Placeholder``s, the object is
a ``Network`` graph (BAISPSPNet.Network) that ``engine.Engine`` lowers, and ``build()`` returns the layer handles of
the four attention logit maps (coarse to fine) and the class logits.

  * trunk: slim/nets/vgg.py:187-196 -- conv1_1 .. conv5_3 (3x3 SAME, bias, ReLU) and the 2x2/2 max-pools;
    block1..4 = conv2_2, conv3_3, conv4_3, conv5_3 (:680-683)
  * the click map is nearest-resized to block4 and multiplies it (:771-774); every attention block (:732-754) is
    3x3 -> BN/ReLU -> 3x3 (32) -> BN/ReLU -> 1x1 (2, bias) -> softmax -> p1 -> where(p1 > 0.9, p1, 0) -> multiply the
    block input -> 1x1 -> nearest resize to the next (finer) block, whose features it gates in turn
  * class head on the coarsest attention output (:711-729, ``p_size`` / ``k_size`` = 15 / 3 in the reference, which
    hard-wires S = 720; here p_size defaults to block4 size // 3)

``width`` scales the trunk's channel counts (1.0 = the reference) so tests can run a narrow copy.
"""
from __future__ import annotations

from .BAISPSPNet import Network, Placeholder, PSPNet  # noqa: F401

VGG_BLOCKS = ((1, 2, 64), (2, 2, 128), (3, 3, 256), (4, 3, 512), (5, 3, 512))


class LinkNet(Network):

    def __init__(self, input_data, input_mask, is_training=True, num_classes=21, width=1.0, p_size=None, k_size=3,
                 thr=0.9):
        self.num_classes, self.width, self.p_size, self.k_size, self.thr = num_classes, width, p_size, k_size, thr
        self.attentions, self.classes, self.segments = [], [], []
        Network.__init__(self, {'data': input_data, 'mask': input_mask}, num_classes, 2, True, is_training)

    def build(self):
        return self.segments, self.attentions, self.classes

    def _attention(self, lvl, source, c_in_size, c_out, out_hw):
        sc = "attention_%d/attention_%d_attention" % (lvl, lvl)
        (self.feed(source)
         .conv(3, 3, c_in_size // 4, 1, 1, biased=False, relu=False, padding='SAME', name=sc + '/a_conv_1')
         .batch_normalization(relu=True, name=sc + '/a_conv_1_bn')
         .conv(3, 3, 32, 1, 1, biased=False, relu=False, padding='SAME', name=sc + '/a_conv_2')
         .batch_normalization(relu=True, name=sc + '/a_conv_2_bn')
         .conv(1, 1, 2, 1, 1, biased=True, relu=False, name=sc + '/a_conv_3')
         .softmax_gate(sel=1, thr=self.thr, name=sc + '/gate'))
        (self.feed(source, sc + '/gate')
         .mask_multiply(name=sc + '/mask_multiply')
         .conv(1, 1, c_out, 1, 1, biased=False, relu=False, padding='SAME', name=sc + '/a_conv_o'))
        self.attentions.append(self.layers[sc + '/a_conv_3'])
        (self.feed(sc + '/a_conv_o').resize_nearest(out_hw, name=sc + '/a_conv_o_resize'))
        (self.feed(sc + '/gate').resize_nearest(out_hw, name=sc + '/gate_resize'))
        return sc

    def setup(self, is_training, num_classes, num_segment, last_pool_size, filter_number):
        ch = {}
        self.feed('data')
        for blk, reps, c in VGG_BLOCKS:
            c = max(8, int(c * self.width))
            for r in range(1, reps + 1):
                self.conv(3, 3, c, 1, 1, biased=True, relu=True, padding='SAME',
                          name='vgg_16/conv%d/conv%d_%d' % (blk, blk, r))
            ch[blk] = c
            if blk < 5:
                self.max_pool(2, 2, 2, 2, name='vgg_16/pool%d' % blk)
        blocks = {1: 'vgg_16/conv2/conv2_2', 2: 'vgg_16/conv3/conv3_3', 3: 'vgg_16/conv4/conv4_3',
                  4: 'vgg_16/conv5/conv5_3'}
        c1, c2, c3, c4 = ch[2], ch[3], ch[4], ch[5]
        hw = {k: self.layers[v].shape[:2] for k, v in blocks.items()}
        # initial attention = the click map at block4 resolution; Net.concat([f, f]) of the reference is two
        # (identical) gated copies here -- a concat slice has one producer
        self.feed('mask').resize_nearest(hw[4], name='mask_block4')
        self.feed(blocks[4], 'mask_block4').mask_multiply(name='block4_attention_multiply_add')
        self.feed(blocks[4], 'mask_block4').mask_multiply(name='block4_attention_multiply_add_b')
        (self.feed('block4_attention_multiply_add', 'block4_attention_multiply_add_b')
         .concat(axis=-1, name='attention_4_concat'))
        sc = self._attention(4, 'attention_4_concat', c4, c3, hw[3])
        # class decoder on the coarsest attention output
        dc = "attention_4/segment_attention_4_decoder"
        p_size = self.p_size if self.p_size is not None else hw[4][0] // 3
        (self.feed(sc + '/a_conv_o')
         .conv(3, 3, c4, 1, 1, biased=False, relu=False, padding='SAME', name=dc + '/d_c_conv_1')
         .batch_normalization(relu=True, name=dc + '/d_c_conv_1_bn')
         .conv(3, 3, c4 * 2, 1, 1, biased=False, relu=False, padding='SAME', name=dc + '/d_c_conv_2')
         .batch_normalization(relu=True, name=dc + '/d_c_conv_2_bn')
         .avg_pool(p_size, p_size, p_size, p_size, name=dc + '/class_attention_pool')
         .conv(self.k_size, self.k_size, c4 * 4, self.k_size, self.k_size, name=dc + '/class_attention_conv')
         .squeeze(name=dc + '/class_attention_squeeze')
         .fc(num_out=num_classes, relu=False, name=dc + '/class_attention_fc'))
        self.classes.append(self.layers[dc + '/class_attention_fc'])
        for lvl, c_blk, c_out, nxt in ((3, c3, c2, 2), (2, c2, c1, 1), (1, c1, c1, 1)):
            (self.feed(blocks[lvl], sc + '/gate_resize')
             .mask_multiply(name='block%d_attention_multiply_add' % lvl))
            (self.feed('block%d_attention_multiply_add' % lvl, sc + '/a_conv_o_resize')
             .concat(axis=-1, name='attention_%d_concat' % lvl))
            sc = self._attention(lvl, 'attention_%d_concat' % lvl, c_blk, c_out, hw[nxt])


class LinkNetTop(Network):
    """The current-HEAD network (BAISNet.py:117-269: ``LinkNet(input_data, is_training, num_classes).build()`` ->
    ``(segments, features)``): vgg_16 trunk, per level a deep-supervised side head
    (1x1 -> nearest resize -> 3x3 -> 1x1 -> 3x3 to 2 logits) and a decoder (1x1 -> nearest resize -> 3x3 -> 1x1), one
    residual add at block1, a final 3x3 head; every convolution has bias (+ReLU except the logit heads), no BN.
    Labels are full-resolution {0,1} maps (label_stride 1); the loss is ``cal_loss`` of BAISRunnerTrain.py:97-115."""

    label_stride = 1

    def __init__(self, input_data, is_training=True, num_classes=21, width=1.0):
        self.width = width
        self.attentions, self.classes, self.segments = [], [], []     # (attentions: the engine's name for the 2-ch heads)
        Network.__init__(self, {'data': input_data}, num_classes, 2, True, is_training)

    def build(self):
        return self.segments, {"segment": self.decoder_outputs}

    def _decoder(self, source, scope, c_in, c_out, out_hw, head):
        (self.feed(source)
         .conv(1, 1, c_in // 4, 1, 1, biased=True, relu=True, padding='SAME', name=scope + '/d_s_conv_1')
         .resize_nearest(out_hw, name=scope + '/resize')
         .conv(3, 3, c_in // 4, 1, 1, biased=True, relu=True, padding='SAME', name=scope + '/d_s_conv_2')
         .conv(1, 1, c_out, 1, 1, biased=True, relu=True, name=scope + '/d_s_conv_3'))
        if head:
            self.conv(3, 3, 2, 1, 1, biased=True, relu=False, name=scope + '/d_s_conv_4')     # (VALID: the map shrinks by 2)
            return scope + '/d_s_conv_4'
        return scope + '/d_s_conv_3'

    def setup(self, is_training, num_classes, num_segment, last_pool_size, filter_number):
        ch = {}
        self.feed('data')
        for blk, reps, c in VGG_BLOCKS:
            c = max(8, int(c * self.width))
            for r in range(1, reps + 1):
                self.conv(3, 3, c, 1, 1, biased=True, relu=True, padding='SAME',
                          name='vgg_16/conv%d/conv%d_%d' % (blk, blk, r))
            ch[blk] = c
            if blk < 5:
                self.max_pool(2, 2, 2, 2, name='vgg_16/pool%d' % blk)
        blocks = {1: 'vgg_16/conv2/conv2_2', 2: 'vgg_16/conv3/conv3_3', 3: 'vgg_16/conv4/conv4_3',
                  4: 'vgg_16/conv5/conv5_3'}
        c = {1: ch[2], 2: ch[3], 3: ch[4], 4: ch[5]}
        hw = {k: self.layers[v].shape[:2] for k, v in blocks.items()}
        self.decoder_outputs = []
        cur = blocks[4]
        for lvl, nxt in ((4, 3), (3, 2), (2, 1), (1, 1)):
            if lvl == 1:
                self.feed(blocks[1], cur).add(name='attention_1/add')
                cur = 'attention_1/add'
            head = self._decoder(cur, "attention_%d/segment_side_%d" % (lvl, lvl), c[lvl], c[nxt], hw[nxt], True)
            self.attentions.append(self.layers[head])
            cur = self._decoder(cur, "attention_%d/%d" % (lvl, lvl), c[lvl], c[nxt], hw[nxt], False)
            self.decoder_outputs.append(self.layers[cur])
        self.feed(cur).conv(3, 3, 2, 1, 1, biased=True, relu=False, name='attention_0')
        self.attentions.append(self.layers['attention_0'])
        self.segments = list(self.attentions)


class BAISNet(PSPNet):
    """Cascaded attention re-decoding: back/8AttentionU/BAISNet.py:13-650,
    ``BAISNet(input_data, is_training, num_classes, num_segment, segment_attention, last_pool_size, filter_number,
    attention_module_num).build()`` -> ``(segments, attentions, classes)``.

    The 2AddClass trunk, then four pyramid decoders (:487-530) in the scopes '', attention_1, attention_2,
    attention_3; the last ``attention_module_num`` decode 2 channels, the others ``num_segment``.  Decoder k reads the
    feature multiplied by the softmax attention channel of decoder k-1 (:584-637), every gated feature also feeds a
    class head (:533-544).  ``segments`` are the SIGMOID outputs -- ``cal_loss`` consumes them as logits."""

    cascade = True
    SCOPES = ('', 'attention_1/', 'attention_2/', 'attention_3/')

    def __init__(self, input_data, is_training=True, num_classes=21, num_segment=4, segment_attention=1,
                 last_pool_size=90, filter_number=32, attention_module_num=2):
        self.variant = "4BorderClass"
        self.attention_class = None
        self.num_classes, self.num_segment = num_classes, num_segment
        self.segment_attention, self.attention_module_num = segment_attention, attention_module_num
        self.last_pool_size, self.filter_number = last_pool_size, filter_number
        self.segments, self.attentions, self.classes = [], [], []
        self.segment_logits = []
        Network.__init__(self, {'data': input_data}, num_classes, num_segment, True, is_training, last_pool_size,
                         filter_number)

    def build(self):
        return self.segments, self.attentions, self.classes

    def _decoder(self, source, scope, nseg):
        F, P = self.filter_number, self.last_pool_size
        lg = self._pyramid_decoder(source, F, P, nseg, 'conv6_n_4', scope)
        self.feed(lg).sigmoid(name=lg + '/sigmoid')
        self.segment_logits.append(self.layers[lg])
        self.segments.append(self.layers[lg + '/sigmoid'])
        return lg

    def _classifies(self, source, scope):
        F, P = self.filter_number, self.last_pool_size
        pool_ratio = 5
        ps = P // pool_ratio
        (self.feed(source)
         .avg_pool(ps, ps, ps, ps, name=scope + 'class_attention_pool')
         .conv(pool_ratio, pool_ratio, F * 16, pool_ratio, pool_ratio, name=scope + 'class_attention_conv')
         .squeeze(name=scope + 'class_attention_squeeze')
         .fc(num_out=self.num_classes, relu=False, name=scope + 'class_attention_fc'))
        self.classes.append(self.layers[scope + 'class_attention_fc'])

    def setup(self, is_training, num_classes, num_segment, last_pool_size, filter_number):
        feat = self._trunk(filter_number, wiring="8AttentionU")
        n = len(self.SCOPES)
        first_two = n - self.attention_module_num            # decoders with index >= first_two have 2 channels
        lg = self._decoder(feat, '', num_segment if first_two > 0 else 2)
        for i in range(1, n + 1):
            sc = self.SCOPES[i] if i < n else ''
            sel = self.segment_attention if (i - 1) < first_two else 1
            gate = lg + '/softmax_attention'
            self.feed(lg).softmax_gate(sel=sel, thr=-1.0, name=gate)       # tf.split(softmax, n)[sel], no threshold
            self.attentions.append(self.layers[gate])
            (self.feed(feat, gate).mask_multiply(name=sc + 'class_attention_multiply'))
            feat = sc + 'class_attention_multiply'                         # (the reference's "multiply * 1")
            self._classifies(feat, sc)
            if i < n:
                lg = self._decoder(feat, sc, num_segment if i < first_two else 2)
