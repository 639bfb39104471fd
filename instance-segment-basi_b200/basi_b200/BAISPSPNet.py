"""PSPNet graph builder with the reference's fluent ``Network`` interface.

Mirrors the call surface of back/2AddClass/BAISPSPNet.py (``layer`` decorator :5-29,
``Network`` :32-255, ``PSPNet.setup`` :258-734) and the head variants of
back/4BorderClass/BAISPSPNet.py:716-737 and back/5COCO/BAISPSPNet.py:716-737.

Where the reference's ``Network`` methods append TensorFlow ops to a tf.Graph, these
append nodes to a small static graph (``Node``) that ``engine.Engine`` lowers to fused
sm_100a kernels behind the C ABI.  ``net.layers[name]`` therefore holds ``Node`` handles
(the analogue of symbolic tf.Tensors): pass them to ``Engine.fetch`` / ``Session.run``.
"""
from __future__ import annotations

import numpy as np


class Node(object):
    """One reference-level op: the analogue of a symbolic tf.Tensor."""
    __slots__ = ("op", "name", "inputs", "attrs", "shape", "index")

    def __init__(self, op, name, inputs, attrs, shape, index):
        self.op, self.name, self.inputs, self.attrs, self.shape, self.index = op, name, inputs, attrs, shape, index

    def get_shape(self):
        return [None] + list(self.shape)

    def __repr__(self):
        return "Node(%s %s %s)" % (self.op, self.name, self.shape)


class Placeholder(Node):
    """Stand-in for tf.placeholder(dtype=float32, shape=(None, H, W, C))."""

    def __init__(self, shape, name="data"):
        shape = tuple(shape)[-3:]
        Node.__init__(self, "data", name, [], {}, shape, -1)


def layer(op):
    """Decorator for composable network layers (same contract as the reference's)."""

    def layer_decorated(self, *args, **kwargs):
        name = kwargs.setdefault('name', self.get_unique_name(op.__name__))
        if len(self.terminals) == 0:
            raise RuntimeError('No input variables found for layer %s.' % name)
        elif len(self.terminals) == 1:
            layer_input = self.terminals[0]
        else:
            layer_input = list(self.terminals)
        layer_output = op(self, layer_input, *args, **kwargs)
        self.layers[name] = layer_output
        self.feed(layer_output)
        return self

    return layer_decorated


def _same_out(n, s):
    return -(-n // s)


class Network(object):

    def __init__(self, inputs, num_classes, num_segment, trainable=True, is_training=False, last_pool_size=90,
                 filter_number=64, **extra):
        self.inputs = inputs
        self.terminals = []
        self.layers = dict(inputs)
        self.trainable = trainable
        self.is_training = is_training
        self.nodes = []        # creation order == topological order
        self.variables = {}    # tf variable name -> shape (insertion ordered)
        for v in inputs.values():
            self._register(v)
        self.extra = extra
        self.setup(is_training, num_classes, num_segment, last_pool_size, filter_number)

    def setup(self, is_training, num_classes, num_segment, last_pool_min_size, filter_number):
        raise NotImplementedError('Must be implemented by the subclass.')

    # ---- graph bookkeeping ----
    def _register(self, node):
        node.index = len(self.nodes)
        self.nodes.append(node)
        return node

    def _node(self, op, name, inputs, shape, **attrs):
        return self._register(Node(op, name, list(inputs), attrs, tuple(shape), -1))

    def feed(self, *args):
        assert len(args) != 0
        self.terminals = []
        for fed_layer in args:
            if isinstance(fed_layer, str):
                try:
                    fed_layer = self.layers[fed_layer]
                except KeyError:
                    raise KeyError('Unknown layer name fed: %s' % fed_layer)
            self.terminals.append(fed_layer)
        return self

    def get_output(self):
        return self.terminals[-1]

    def get_unique_name(self, prefix):
        index = sum(t.startswith(prefix) for t, _ in self.layers.items()) + 1
        return '%s_%d' % (prefix, index)

    def make_var(self, name, shape):
        """Registers a variable under its TF scope name; the engine owns the storage."""
        if name in self.variables:
            raise ValueError("variable %s already exists" % name)
        self.variables[name] = tuple(int(s) for s in shape)
        return name

    def consumers(self, node):
        return [n for n in self.nodes if any(i is node for i in n.inputs)]

    # ---- layers ----
    @layer
    def zero_padding(self, input, paddings, name):
        h, w, c = input.shape
        return self._node("zero_padding", name, [input], (h + 2 * paddings, w + 2 * paddings, c), pad=paddings)

    def _conv_common(self, input, k_h, k_w, c_o, stride, dilation, name, relu, padding, biased):
        h, w, c = input.shape
        wname = self.make_var(name + '/weights', [k_h, k_w, c, c_o])
        bname = self.make_var(name + '/biases', [c_o]) if biased else None
        if padding == "SAME":
            oh, ow = _same_out(h, stride), _same_out(w, stride)
        else:
            oh = (h - ((k_h - 1) * dilation + 1)) // stride + 1
            ow = (w - ((k_w - 1) * dilation + 1)) // stride + 1
        if oh <= 0 or ow <= 0:
            raise ValueError("conv %s: input %s too small" % (name, (h, w)))
        return self._node("conv", name, [input], (oh, ow, c_o), k_h=k_h, k_w=k_w, stride=stride, dilation=dilation,
                          relu=relu, padding=padding, weights=wname, biases=bname)

    @layer
    def conv(self, input, k_h, k_w, c_o, s_h, s_w, name, relu=True, padding="VALID", group=1, biased=True):
        assert s_h == s_w and group == 1
        return self._conv_common(input, k_h, k_w, c_o, s_h, 1, name, relu, padding, biased)

    @layer
    def atrous_conv(self, input, k_h, k_w, c_o, dilation, name, relu=True, padding="VALID", group=1, biased=True):
        assert group == 1
        return self._conv_common(input, k_h, k_w, c_o, 1, dilation, name, relu, padding, biased)

    @layer
    def relu(self, input, name):
        return self._node("relu", name, [input], input.shape)

    @layer
    def max_pool(self, input, k_h, k_w, s_h, s_w, name, padding="VALID"):
        h, w, c = input.shape
        if k_h == 2 and k_w == 2 and s_h == 2 and s_w == 2 and padding == "VALID":
            return self._node("max_pool", name, [input], (h // 2, w // 2, c), k=2)     # slim vgg_16 pools (variant B)
        if not (k_h == 3 and k_w == 3 and s_h == 2 and s_w == 2 and padding == "SAME"):
            raise NotImplementedError("max_pool: only 3x3 s2 SAME and 2x2 s2 VALID are on the hot path")
        return self._node("max_pool", name, [input], (_same_out(h, 2), _same_out(w, 2), c), k=3)

    @layer
    def resize_nearest(self, input, size, name):
        """tf.image.resize_nearest_neighbor (TF1 default)."""
        return self._node("resize_nearest", name, [input], (int(size[0]), int(size[1]), input.shape[2]))

    @layer
    def mask_multiply(self, inputs, name):
        """tf.multiply(feature, mask): a 1-channel float32 map broadcast over the feature channels."""
        feature, mask = inputs
        assert mask.shape[2] == 1 and mask.shape[:2] == feature.shape[:2], (feature.shape, mask.shape)
        return self._node("mask_multiply", name, [feature, mask], feature.shape)

    @layer
    def softmax_gate(self, input, name, sel=1, thr=0.9):
        """softmax over the channels -> channel `sel` -> tf.where(p > thr, p, 0) (LinkNet._attention)."""
        h, w, c = input.shape
        return self._node("softmax_gate", name, [input], (h, w, 1), sel=sel, thr=thr)

    @layer
    def sigmoid(self, input, name):
        """tf.nn.sigmoid on a float32 logit map (back/8AttentionU/BAISNet.py:527)."""
        return self._node("sigmoid", name, [input], input.shape)

    @layer
    def avg_pool(self, input, k_h, k_w, s_h, s_w, name, padding="VALID"):
        if not (k_h == k_w == s_h == s_w and padding == "VALID"):
            raise NotImplementedError("avg_pool: only k == s VALID is on the hot path")
        h, w, c = input.shape
        return self._node("avg_pool", name, [input], (h // k_h, w // k_w, c), k=k_h)

    @layer
    def concat(self, inputs, axis, name, dtype=None):
        assert axis in (-1, 3)
        h, w, _ = inputs[0].shape
        assert len(set(id(i) for i in inputs)) == len(inputs), "concat of one tensor with itself: feed two nodes"
        return self._node("concat", name, inputs, (h, w, sum(i.shape[2] for i in inputs)), dtype=dtype)

    @layer
    def add(self, inputs, name):
        return self._node("add", name, inputs, inputs[0].shape)

    @layer
    def fc(self, input, num_out, name, relu=True):
        dim = int(np.prod(input.shape))
        wname = self.make_var(name + '/weights', [dim, num_out])
        bname = self.make_var(name + '/biases', [num_out])
        return self._node("fc", name, [input], (num_out,), relu=relu, weights=wname, biases=bname)

    @layer
    def batch_normalization(self, input, name, scale_offset=True, relu=False):
        c = input.shape[-1]
        # tf.layers.batch_normalization(name=name) inside tf.variable_scope(name): doubled (leaf) scope
        leaf = name.split('/')[-1]
        g = self.make_var('%s/%s/gamma' % (name, leaf), [c])
        b = self.make_var('%s/%s/beta' % (name, leaf), [c])
        return self._node("batch_normalization", name, [input], input.shape, relu=relu, gamma=g, beta=b,
                          momentum=0.95, epsilon=1e-5)

    @layer
    def resize_bilinear(self, input, size, name):
        return self._node("resize_bilinear", name, [input], (int(size[0]), int(size[1]), input.shape[2]))

    @layer
    def multiply(self, inputs, name, num_segment=1, segment_place=0):
        return self._node("multiply", name, inputs, inputs[0].shape, num_segment=num_segment,
                          segment_place=segment_place)

    @layer
    def squeeze(self, inputs, name):
        h, w, c = inputs.shape
        assert h == 1 and w == 1, "squeeze expects a 1x1 map, got %s" % (inputs.shape,)
        return self._node("squeeze", name, [inputs], (c,))


# name of the segment head / class fc per reference snapshot, and which logit channel gates the class head
VARIANTS = {
    "1NoClass": dict(seg="conv6_n", fc=None, place=0),
    "2AddClass": dict(seg="conv6_n", fc="class_attention_fc", place=0),
    "3ThreeClass": dict(seg="conv6_n_3", fc="class_attention_fc", place=1),
    "4BorderClass": dict(seg="conv6_n_4", fc="class_attention_fc", place=1),
    "5COCO": dict(seg="conv6_n_3_coco", fc="class_attention_fc_coco", place=2),
}


class PSPNet(Network):
    """Half-width dilated ResNet-101 + pyramid pooling + segment head + attention-class head.

    PSPNet({'data': placeholder}, is_training, num_classes, num_segment, last_pool_size, filter_number
           [, attention_class][, variant])
    ``variant`` selects the reference snapshot whose head naming is used (default 2AddClass).
    """

    def __init__(self, inputs, num_classes, num_segment, trainable=True, is_training=False, last_pool_size=90,
                 filter_number=64, attention_class=None, variant="2AddClass"):
        if variant not in VARIANTS:
            raise ValueError("unknown variant %r" % (variant,))
        self.variant = variant
        self.attention_class = attention_class
        self.num_classes, self.num_segment = num_classes, num_segment
        self.last_pool_size, self.filter_number = last_pool_size, filter_number
        Network.__init__(self, inputs, num_classes, num_segment, trainable, is_training, last_pool_size,
                         filter_number)

    def _bottleneck(self, prefix, source, mid, stride, dilation, project, sum_source=None):
        """One bottleneck block reading `source` (the previous junction's ReLU).  `sum_source` (8AttentionU wiring
        only) names the previous junction's PRE-ReLU sum: the projection of an entry block / the 1x1_reduce of any
        other block reads that instead (see _trunk)."""
        fn_out = mid * 4
        if project:
            (self.feed(sum_source or source)
             .conv(1, 1, fn_out, stride, stride, biased=False, relu=False, name=prefix + '_1x1_proj')
             .batch_normalization(relu=False, name=prefix + '_1x1_proj_bn'))
            shortcut = prefix + '_1x1_proj_bn'
        else:
            shortcut = source
        (self.feed(source if (project or not sum_source) else sum_source)
         .conv(1, 1, mid, stride, stride, biased=False, relu=False, name=prefix + '_1x1_reduce')
         .batch_normalization(relu=True, name=prefix + '_1x1_reduce_bn')
         .zero_padding(paddings=dilation, name='padding_' + prefix))
        if dilation == 1:
            self.conv(3, 3, mid, 1, 1, biased=False, relu=False, name=prefix + '_3x3')
        else:
            self.atrous_conv(3, 3, mid, dilation, biased=False, relu=False, name=prefix + '_3x3')
        (self.batch_normalization(relu=True, name=prefix + '_3x3_bn')
         .conv(1, 1, fn_out, 1, 1, biased=False, relu=False, name=prefix + '_1x1_increase')
         .batch_normalization(relu=False, name=prefix + '_1x1_increase_bn'))
        (self.feed(shortcut, prefix + '_1x1_increase_bn')
         .add(name=prefix)
         .relu(name=prefix + '/relu'))
        return prefix + '/relu'

    def _trunk(self, F, wiring="2AddClass"):
        """conv1_1 .. conv5_3/relu (2AddClass/BAISPSPNet.py:173-470); returns the name of the last layer.

        wiring="8AttentionU": back/8AttentionU/BAISNet.py:133-480 unrolls the same trunk by hand with one variable
        (`net_input`) that still holds the junction SUM when the next convolution is built (:163-165), so there the
        1x1_reduce of every non-entry block except conv2_2 and the 1x1_proj of conv3_1 / conv4_1 / conv5_1 read the
        pre-ReLU sum of the previous junction; shortcuts, conv2_2_1x1_reduce and the entry blocks' 1x1_reduce read
        its ReLU.  (Found by running that file: tests/golden/reference_net_8AttentionU.json "wiring".)"""
        (self.feed('data')
         .conv(3, 3, F, 2, 2, biased=False, relu=False, padding='SAME', name='conv1_1_3x3_s2_n')
         .batch_normalization(relu=False, name='conv1_1_3x3_s2_bn')
         .relu(name='conv1_1_3x3_s2_bn_relu')
         .conv(3, 3, F, 1, 1, biased=False, relu=False, padding='SAME', name='conv1_2_3x3')
         .batch_normalization(relu=True, name='conv1_2_3x3_bn')
         .conv(3, 3, F * 2, 1, 1, biased=False, relu=False, padding='SAME', name='conv1_3_3x3')
         .batch_normalization(relu=True, name='conv1_3_3x3_bn')
         .max_pool(3, 3, 2, 2, padding='SAME', name='pool1_3x3_s2'))
        cur = 'pool1_3x3_s2'
        for stage, blocks, mult, stride, dilation in ((2, 3, 1, 1, 1), (3, 4, 2, 2, 1), (4, 23, 4, 1, 2),
                                                      (5, 3, 8, 1, 4)):
            for b in range(1, blocks + 1):
                prefix = 'conv%d_%d' % (stage, b)
                prev_sum = cur[:-len('/relu')] if cur.endswith('/relu') else None
                reads_sum = wiring == "8AttentionU" and prev_sum is not None and prefix != 'conv2_2'
                cur = self._bottleneck(prefix, cur, F * mult, stride if b == 1 else 1, dilation, project=(b == 1),
                                       sum_source=prev_sum if reads_sum else None)
        return cur

    def _pyramid_decoder(self, source, F, last_pool_size, num_segment, seg_name, scope=''):
        """Pyramid pooling (levels 1/2/3/6) + conv5_4 + the 1x1 segment head on `source`; layer and variable names are
        prefixed with `scope` (the cascade of back/8AttentionU builds four of these)."""
        shape = self.layers[source].shape[0:2]
        ofn = F * 32 // 4
        for level in (1, 2, 3, 6):
            k = last_pool_size // level
            p = scope + 'conv5_3_pool%d' % level
            (self.feed(source)
             .avg_pool(k, k, k, k, name=p)
             .conv(1, 1, ofn, 1, 1, biased=False, relu=False, name=p + '_conv')
             .batch_normalization(relu=True, name=p + '_conv_bn')
             .resize_bilinear(shape, name=p + '_interp'))
        (self.feed(source, scope + 'conv5_3_pool6_interp', scope + 'conv5_3_pool3_interp',
                   scope + 'conv5_3_pool2_interp', scope + 'conv5_3_pool1_interp')
         .concat(axis=-1, name=scope + 'conv5_3_concat')
         .conv(3, 3, ofn, 1, 1, biased=False, relu=False, padding='SAME', name=scope + 'conv5_4')
         .batch_normalization(relu=True, name=scope + 'conv5_4_bn')
         .conv(1, 1, num_segment, 1, 1, biased=True, relu=False, name=scope + seg_name))
        return scope + seg_name

    def setup(self, is_training, num_classes, num_segment, last_pool_size, filter_number):
        v = VARIANTS[self.variant]
        F = filter_number
        self._trunk(F)
        self._pyramid_decoder('conv5_3/relu', F, last_pool_size, num_segment, v["seg"])
        if v["fc"] is None:
            return
        place = self.attention_class if self.attention_class is not None else v["place"]
        pool_ratio = 5
        pool_size = last_pool_size // pool_ratio
        (self.feed("conv5_3", v["seg"])
         .multiply(num_segment=num_segment, segment_place=place, name="class_attention_multiply")
         .avg_pool(pool_size, pool_size, pool_size, pool_size, name="class_attention_pool")
         .conv(pool_ratio, pool_ratio, F * 16, pool_ratio, pool_ratio, name="class_attention_conv")
         .squeeze(name="class_attention_squeeze")
         .fc(num_out=num_classes, name=v["fc"], relu=False))
