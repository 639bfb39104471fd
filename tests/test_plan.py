"""CPU: the graph builder mirrors the reference topology and the engine lowers it to the expected plan."""
from collections import Counter

import numpy as np
import pytest

from basi_b200.BAISPSPNet import PSPNet, Placeholder, VARIANTS
from basi_b200.engine import Engine
from oracle import basi_oracle as O


def build(variant="2AddClass", nseg=1, S=320, F=32, classes=21):
    return PSPNet({'data': Placeholder((None, S, S, 4))}, num_classes=classes, num_segment=nseg, is_training=True,
                  last_pool_size=S // 8, filter_number=F, variant=variant)


@pytest.mark.parametrize("variant,nseg,classes", [("1NoClass", 1, 21), ("2AddClass", 1, 21), ("4BorderClass", 4, 21),
                                                  ("5COCO", 3, 91)])
def test_variables_match_oracle_inventory(variant, nseg, classes):
    net = build(variant, nseg, classes=classes)
    spec = O.param_specs(variant, classes, nseg, 32)
    assert dict(net.variables) == {k: tuple(v) for k, v in spec.items()}


def test_layer_shapes_at_320():
    net = build()
    L = net.layers
    assert L['conv1_3_3x3_bn'].shape == (160, 160, 64)
    assert L['pool1_3x3_s2'].shape == (80, 80, 64)
    assert L['conv2_3/relu'].shape == (80, 80, 128)
    assert L['conv3_4/relu'].shape == (40, 40, 256)
    assert L['conv4_23/relu'].shape == (40, 40, 512)
    assert L['conv5_3/relu'].shape == (40, 40, 1024)
    assert [L['conv5_3_pool%d' % l].shape[0] for l in (1, 2, 3, 6)] == [1, 2, 3, 6]
    assert L['conv5_3_concat'].shape == (40, 40, 2048)
    assert L['conv6_n'].shape == (40, 40, 1)
    assert L['class_attention_pool'].shape == (5, 5, 1024)
    assert L['class_attention_fc'].shape == (21,)
    assert L['conv6_n'].get_shape()[1:3] == [40, 40]


def test_counts_match_survey():
    net = build()
    ops = Counter(n.op for n in net.nodes)
    assert ops["conv"] + ops["fc"] == 114 and ops["batch_normalization"] == 111 and ops["add"] == 33


def test_lowered_plan_structure():
    e = Engine(build(), 2, "bf16", True, dict(kind="bce", pos_weight=3.0, class_weight=0.2), dry_run=True)
    f, b = Counter(n for n, _, _, _ in e.fwd), Counter(n for n, _, _, _ in e.bwd)
    assert e.n_params == 29571606
    assert f["basi_conv_fprop"] == 112 and f["basi_skinny_fwd_ws"] == 2
    assert f["basi_bn_stats"] == 111 and f["basi_bn_apply"] == 107          # 4 proj BNs folded into junctions
    assert f["basi_bn_finalize"] == 0 and b["basi_bn_bwd_finalize"] == 0    # finalize fused into the reductions
    assert b["basi_conv_wgrad"] == 112 and b["basi_conv_dgrad"] == 111      # conv1_1 needs no dgrad
    assert b["basi_bn_bwd_apply"] == 111
    # the first backward call is the class-head adjoint, the last one the stem's wgrad
    assert e.bwd[-1][0] == "basi_conv_wgrad"


def test_segment_only_plan_has_no_class_head():
    e = Engine(build("1NoClass"), 2, "f32", True, dict(kind="bce", pos_weight=3.0), dry_run=True)
    assert e.cls_logits is None and e.n_params == 16453121
    assert not any(n.startswith("basi_skinny") or n.startswith("basi_gate") for n, _, _, _ in e.fwd + e.bwd)


def test_unknown_variant_and_dry_run_cannot_execute():
    with pytest.raises(ValueError):
        build("nope")
    e = Engine(build(S=64, F=8), 1, "f32", False, None, dry_run=True)
    from basi_b200 import BasiError
    with pytest.raises(BasiError):
        e.forward_device()


def test_click_lut_matches_reference_expression_bitwise():
    from basi_b200.BAISData import Data, click_lut
    lut = click_lut((320, 320), 30)
    assert lut.size == 2 * 319 * 319 + 1
    for where in ([137, 201], [0, 0], [319, 319], [160, 5]):
        ref = O.mask_gaussian((320, 320), where, 30)
        yy, xx = np.mgrid[0:320, 0:320]
        d2 = (xx - where[1]) ** 2 + (yy - where[0]) ** 2
        assert np.array_equal(lut[d2].view(np.uint32), ref.view(np.uint32))
        assert np.array_equal(Data._mask_gaussian((320, 320), where).view(np.uint32), ref.view(np.uint32))
    lut20 = click_lut((64, 48), 20)
    ref = O.mask_gaussian((64, 48), [10, 40], 20)
    yy, xx = np.mgrid[0:64, 0:48]
    assert np.array_equal(lut20[(xx - 40) ** 2 + (yy - 10) ** 2].view(np.uint32), ref.view(np.uint32))


def test_synthetic_batches_are_reference_shaped():
    from basi_b200.BAISData import SyntheticData
    s = SyntheticData(3, (64, 64), 8, 21, 4, seed=1)
    img, clicks, lab, cls = s.next_batch()
    assert img.shape == (3, 64, 64, 3) and img.dtype == np.uint8
    assert lab.shape == (3, 8, 8, 1) and set(np.unique(lab)) <= {0, 1, 2, 3}
    for b in range(3):
        assert lab[b, clicks[b, 0] // 8, clicks[b, 1] // 8, 0] == 1       # the click lies on the instance
    data, ann, c, raw, mask = SyntheticData(2, (64, 64), 8, 21, 1, seed=1).next_batch_train()
    assert data[0].shape == (64, 64, 4) and data[0].dtype == np.float32 and ann[0].shape == (8, 8, 1)


def test_psp_branches_are_tagged_and_their_pool_adjoints_deferred():
    """The four pyramid branches run on their own streams: every call of a branch carries its tag, the shared-tensor
    read-modify-write (avgpool adjoint) carries none and is emitted after the last branch call."""
    eng = Engine(build("1NoClass", S=64, F=8), 2, precision="bf16", dry_run=True,       # (f32: four order-free pools)
                 loss=dict(kind="bce", pos_weight=3.0))
    for lst in (eng.fwd, eng.bwd):
        tags = [m.get("branch") for _, _, _, m in lst]
        assert set(t for t in tags if t is not None) == {0, 1, 2, 3}
        idx = [i for i, t in enumerate(tags) if t is not None]
        # one contiguous fork/join region per pass
        assert all(tags[i] is not None for i in range(idx[0], idx[-1] + 1))
    names = [n for n, _, _, _ in eng.bwd]
    tags = [m.get("branch") for _, _, _, m in eng.bwd]
    # the four pools are one fused pass (forward: before the fork; backward: one read-modify-write after the join)
    pools = [i for i, n in enumerate(names) if n == "basi_avgpool_multi_bwd"]
    last_branch = max(i for i, t in enumerate(tags) if t is not None)
    assert len(pools) == 1 and tags[pools[0]] is None and pools[0] == last_branch + 1
    assert "basi_avgpool_bwd" not in names
    # ... and before the consumer of the tensor it adds into (conv5_3's junction backward)
    assert names[pools[0] + 1].startswith("basi_bn_bwd"), names[pools[0] + 1]
    fnames = [n for n, _, _, _ in eng.fwd]
    ftags = [m.get("branch") for _, _, _, m in eng.fwd]
    fp = fnames.index("basi_avgpool_multi_fwd")
    assert ftags[fp] is None and fp + 1 == min(i for i, t in enumerate(ftags) if t is not None)


# ---------------------------------------------------------------------------------------------------------------
# round 2: host logic of the new paths (all CPU, dry-run plans)
# ---------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("B", [16, 64, 128])
def test_class_head_lowers_to_skinny_gemms_at_any_batch(B):
    """the skinny-GEMM lowering used to accept batch sizes its kernels rejected (B = 49..64, B > 64).  One
    predicate (basi_skinny_supported) now gates both lowerings and the kernels chunk the rows."""
    from basi_b200 import _lib
    lib = _lib.load()
    assert lib.basi_skinny_supported(B, 25 * 1024, 512) == 1 and lib.basi_skinny_supported(B, 512, 21) == 1
    e = Engine(build(), B, "bf16", True, dict(kind="bce", pos_weight=3.0, class_weight=0.2), dry_run=True)
    f = Counter(n for n, _, _, _ in e.fwd)
    assert f["basi_skinny_fwd_ws"] == 2


def test_experiment_switches_need_the_master_switch(monkeypatch, capsys):
    from basi_b200 import engine as E
    monkeypatch.delenv("BASI_EXPERIMENTS", raising=False)
    monkeypatch.setenv("BASI_NO_MASK_BITS", "1")
    E._WARNED.discard("BASI_NO_MASK_BITS")
    assert E._exp_env("BASI_NO_MASK_BITS") is None
    assert "ignored" in capsys.readouterr().err
    monkeypatch.setenv("BASI_EXPERIMENTS", "1")
    assert E._exp_env("BASI_NO_MASK_BITS") == "1"


def test_variant_b_plan_and_inventory():
    from basi_b200.BAISNet import LinkNet
    net = LinkNet(Placeholder((None, 320, 320, 3)), Placeholder((None, 320, 320, 1), name="mask"), num_classes=21)
    spec = O.linknet_b_specs(21, 1.0)
    assert list(net.variables.items()) == [(k, tuple(v)) for k, v in spec.items()]
    segs, atts, classes = net.build()
    assert [a.shape for a in atts] == [(20, 20, 2), (40, 40, 2), (80, 80, 2), (160, 160, 2)] and classes[0].shape == (21,)
    e = Engine(net, 2, "f32", True, dict(kind="linknet_b"), dry_run=True)
    f, b = Counter(n for n, _, _, _ in e.fwd), Counter(n for n, _, _, _ in e.bwd)
    assert f["basi_maxpool2s2_fwd"] == 4 and f["basi_softmax_gate_fwd"] == 4 and f["basi_mask_mul_fwd"] == 9
    assert f["basi_resize_nearest_fwd"] == 9
    # the finest attention output and both of its resized copies feed nothing: no adjoint is emitted for them
    assert b["basi_softmax_gate_bwd"] == 3 and b["basi_resize_nearest_bwd"] == 6 and b["basi_mask_mul_bwd"] == 8
    losses = Counter(n for n, _, _, _ in e.lossl)
    assert losses["basi_wbce_fwd_bwd"] == 4 and losses["basi_onehot2_f32"] == 4 and losses["basi_softmax_ce_fwd_bwd"] == 1


def test_top_level_linknet_plan_and_inventory():
    from basi_b200.BAISNet import LinkNetTop
    net = LinkNetTop(Placeholder((None, 320, 320, 3)))
    spec = O.linknet_top_specs(1.0)
    assert list(net.variables.items()) == [(k, tuple(v)) for k, v in spec.items()]
    segs, feats = net.build()
    assert [s.shape for s in segs] == [(38, 38, 2), (78, 78, 2), (158, 158, 2), (158, 158, 2), (158, 158, 2)]
    e = Engine(net, 2, "f32", True, dict(kind="linknet_b", pos_weight=1.0), dry_run=True)
    f, b = Counter(n for n, _, _, _ in e.fwd), Counter(n for n, _, _, _ in e.bwd)
    assert f["basi_add_fwd"] == 1 and b["basi_add_bwd"] == 1 and f["basi_resize_nearest_fwd"] == 8
    assert e.label_seg.shape == (2, 320, 320, 1)                  # full-resolution labels (BAISRunnerTrain.py:38)


def test_set_trainable_ranges_cover_exactly_the_named_variables():
    e = Engine(build(), 2, "bf16", True, dict(kind="bce", pos_weight=3.0, class_weight=0.2), dry_run=True)
    n = e.set_trainable("class_attention")
    assert n == 1                                                  # the class head's variables are contiguous
    off, length = e._train_ranges[0]
    names = [k for k in e.param_index if "class_attention" in k]
    assert off == e.param_index[names[0]][0] and off + length == e.n_flat
    with pytest.raises(KeyError):
        e.set_trainable("no_such_scope")
    assert e.set_trainable(None) == 0 and e._train_ranges is None


def test_label_encodings_known_answers():
    """uint8 semantics of back/4BorderClass/BAISData.py:143-160 and back/5COCO/BAISData.py:361-369."""
    ann = np.array([[0, 1, 2, 255, 85, 86, 170, 254]], dtype=np.uint8)
    # has_255=True, attended instance 2: background (0) wraps to 255 // 84 = 3, border 255 -> 171 -> 2, instance 2 -> 85
    # -> 1, other instances: (k - 1) // 84  (ids >= 85 alias: 85 -> 1, 86 -> 1, 170 -> 2, 254 -> 3: reference quirk)
    assert O.encode_labels_border(ann, 2).tolist() == [[3, 0, 1, 2, 1, 1, 2, 3]]
    # has_255=False: border -> background -> (255) // 127 = 2, instance 2 -> 128 -> 1, others (k - 1) // 127
    assert O.encode_labels_three(ann, 2).tolist() == [[2, 0, 1, 2, 0, 0, 1, 1]]
    assert O.encode_labels_binary(ann, 2).tolist() == [[0, 0, 1, 0, 0, 0, 0, 0]]
    s = np.array([[0, 1, 3, 0]], dtype=np.uint8)
    a = np.array([[0, 0, 1, 1]], dtype=np.uint8)
    assert O.encode_labels_coco(s, a).tolist() == [[0, 1, 2, 2]]
    lab = np.array([[0, 1, 0], [1, 1, 0]])
    assert O.sample_click(lab, 0, 8) == [0, 8] and O.sample_click(lab, 2, 8) == [8, 8]


def test_data_reader_shards_annotations_by_rank():
    from basi_b200.BAISData import Data
    readers = []
    for rank in range(2):
        d = Data.__new__(Data)                                     # the sharding logic without a dataset on disk
        d.rank, d.world, d._epoch, d._seed, d.batch_size = rank, 2, 0, 7, 2
        d._annotations = list(range(10))
        d._order = list(range(10))
        d._random_index = d._order[rank::2]
        d.number_patch = len(d._random_index) // d.batch_size
        d._now = 0
        readers.append(d)
    assert readers[0]._random_index == [0, 2, 4, 6, 8] and readers[1]._random_index == [1, 3, 5, 7, 9]
    for d in readers:
        d._reshuffle()
    a, b = readers[0]._random_index, readers[1]._random_index
    assert sorted(a + b) == list(range(10)) and not set(a) & set(b)     # same permutation, disjoint slices


def test_cascade_plan_and_inventory():
    """F4 (back/8AttentionU): four decoders in the scopes '', attention_1..3, four class heads, variables named like
    the reference's; the plan has one fused pyramid-pool pass per decoder and the cal_loss kernels."""
    from basi_b200.BAISNet import BAISNet
    S, F = 320, 32
    net = BAISNet(Placeholder((None, S, S, 4)), is_training=True, num_classes=21, num_segment=4, segment_attention=1,
                  last_pool_size=S // 8, filter_number=F, attention_module_num=2)
    segs, atts, clss = net.build()
    assert [s.shape for s in segs] == [(40, 40, 4), (40, 40, 4), (40, 40, 2), (40, 40, 2)]
    assert [a.shape for a in atts] == [(40, 40, 1)] * 4 and [c.shape for c in clss] == [(21,)] * 4
    spec = O.attention_u_specs(21, 4, F, 2)
    assert list(net.variables.keys()) == list(spec.keys())
    assert dict(net.variables) == {k: tuple(v) for k, v in spec.items()}
    assert "attention_2/conv5_4_bn/conv5_4_bn/gamma" in net.variables and "class_attention_fc/weights" in net.variables
    e = Engine(net, 2, "bf16", True, dict(kind="cascade"), dry_run=True)
    fwd = Counter(c[0] for c in e.fwd)
    assert fwd["basi_avgpool_multi_fwd"] == 4 and fwd["basi_sigmoid_fwd"] == 4 and fwd["basi_softmax_gate_fwd"] == 4
    loss = Counter(c[0] for c in e.lossl)
    assert loss == {"basi_softmax_ce_fwd_bwd": 6, "basi_wbce_sel_fwd_bwd": 2}
    bwd = Counter(c[0] for c in e.bwd)
    assert bwd["basi_sigmoid_bwd"] == 4 and bwd["basi_softmax_gate_bwd"] == 4 and bwd["basi_mask_mul_bwd"] == 4


def test_cascade_oracle_known_answers():
    """cal_loss of the cascade on hand-checkable inputs: uniform sigmoid outputs 0.5 give CE = log(4) for the
    4-channel maps and 2 * mean(wbce(0.5)) for the 2-channel ones; loss = mean + 0.1 * mean class CE."""
    import torch
    B, P = 1, 2
    segs = [torch.full((B, P, P, 4), 0.5, dtype=torch.float64)] * 2 + [torch.full((B, P, P, 2), 0.5, dtype=torch.float64)] * 2
    cls = [torch.zeros((B, 21), dtype=torch.float64)] * 4
    ls = torch.zeros((B, P, P, 1), dtype=torch.int64)
    la = torch.tensor([1.0, 0.0, 0.0, 1.0], dtype=torch.float64).reshape(B, P, P, 1)
    loss, lseg, lcls = O.attention_u_losses(segs, cls, ls, la, torch.zeros((B,), dtype=torch.int64))
    sp = np.log1p(np.exp(-0.5))                    # softplus(-0.5)
    w1 = 3 * sp                                    # z = 1: 0 * x + 3 * softplus(-x)
    w0 = 0.5 + sp                                  # z = 0: x + softplus(-x)
    expect_seg = (2 * np.log(4.0) + 2 * 2 * (w1 + w0) / 2) / 4
    assert abs(float(lseg) - expect_seg) < 1e-12
    assert abs(float(lcls) - np.log(21.0)) < 1e-12
    assert abs(float(loss) - (expect_seg + 0.1 * np.log(21.0))) < 1e-12
