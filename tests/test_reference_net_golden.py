"""The oracle and the product's graph builder against the REFERENCE's own network / training code.

tests/golden/reference_net_<snapshot>.{npz,json} were produced by tests/golden/make_reference_net_golden.py, which
imports the unmodified reference modules (back/<snapshot>/BAISRunnerTrain.py, BAISPSPNet.py) in the build container
over an eager float64 stand-in for the ~60 `tf.*` calls they make and runs the reference's own `Train.build_net()`.
Held to those files here (CPU, no reference needed at test time):

  * oracle/basi_oracle.py: parameter inventory (names, shapes, creation order), every layer it exposes, logits,
    predictions, losses, learning rate, every gradient and the SGD update -- to float64 round-off
  * the product's host side: `basi_b200.BAISPSPNet.PSPNet` registers the same layers (name, shape, order), the same
    variables and issues the same primitive-op sequence as the reference's builder; `BAISRunnerTrain.SNAPSHOT` holds
    the loss weights / learning-rate constants the reference's build_net used; `Engine.set_trainable`'s class-only
    subset equals the `var_list` of the reference's `train_classes_op`

The -m gpu counterpart (CUDA f32 path against the same files) is tests/test_gpu_net.py::test_train_step_matches_reference_code.
"""
import json
import os
import sys

import numpy as np
import pytest
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))

from oracle import basi_oracle as O                       # noqa: E402
from ref_net_common import kind_of, param_value, summary  # noqa: E402

SNAPSHOTS = ("1NoClass", "2AddClass", "3ThreeClass", "4BorderClass", "5COCO")


def load(snapshot):
    base = os.path.join(HERE, "golden", "reference_net_%s" % snapshot)
    with open(base + ".json") as f:
        meta = json.load(f)
    return meta, np.load(base + ".npz")


def reference_params(meta):
    """{tf name: float32 array} for every TRAINABLE variable the reference's code created."""
    return {n: param_value(n, s, kind_of(n)) for n, s, trainable in meta["variables"] if trainable}


def close(a, b, tol=1e-9):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    assert a.shape == b.shape, (a.shape, b.shape)
    scale = max(float(np.abs(b).max()) if b.size else 0.0, 1e-30)
    err = float(np.abs(a - b).max()) / scale if b.size else 0.0
    assert err <= tol, "rel-max error %.3e > %.1e" % (err, tol)
    return err


@pytest.mark.skipif(not os.path.isdir("/root/reference"), reason="the reference checkout exists in the build container only")
@pytest.mark.parametrize("name", ["2AddClass", "8AttentionU"])
def test_committed_fixture_is_what_the_reference_code_produces_today(name, tmp_path):
    """Provenance: re-run the generator (which imports the reference's modules from /root/reference and executes them
    over the TensorFlow stand-in) and compare with the committed fixture, array for array."""
    import subprocess
    env = dict(os.environ, BASI_GOLDEN_OUT=str(tmp_path))
    subprocess.check_call([sys.executable, os.path.join(HERE, "golden", "make_reference_net_golden.py"), name],
                          env=env, stdout=subprocess.DEVNULL)
    meta, z = load(name)
    with open(str(tmp_path / ("reference_net_%s.json" % name))) as f:
        assert json.load(f) == meta
    fresh = np.load(str(tmp_path / ("reference_net_%s.npz" % name)))
    assert sorted(fresh.files) == sorted(z.files)
    for k in z.files:
        assert np.array_equal(fresh[k], z[k]), k


def product_snapshot_table():
    from basi_b200.BAISRunnerTrain import SNAPSHOT
    return SNAPSHOT


# --------------------------------------------------------------------------------------------------------------
# oracle <-> reference code
# --------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("snapshot", SNAPSHOTS)
def test_oracle_parameter_inventory_is_the_reference_codes(snapshot):
    meta, _ = load(snapshot)
    cfg = meta["config"]
    specs = O.param_specs(snapshot, cfg["num_classes"], cfg.get("num_segment", 1), cfg["filter_number"])
    ref = [(n, tuple(s)) for n, s, trainable in meta["variables"] if trainable]
    assert [(n, tuple(s)) for n, s in specs.items()] == ref          # names, shapes AND creation order
    # what the reference's code creates besides: the moving statistics of every tf.layers.batch_normalization
    rest = [n for n, _, trainable in meta["variables"] if not trainable]
    assert all(n.endswith(("/moving_mean", "/moving_variance")) for n in rest)
    assert len(rest) == 2 * sum(1 for n, _ in ref if n.endswith("/gamma"))
    # minimize(loss) touches every trainable variable; minimize(loss, var_list=...) the class head only
    assert meta["train_op_vars"] == [n for n, _ in ref]
    assert meta["train_classes_op_vars"] == [n for n, _ in ref if "class_attention" in n]


@pytest.mark.parametrize("snapshot", SNAPSHOTS)
def test_oracle_train_step_reproduces_the_reference_code(snapshot):
    meta, z = load(snapshot)
    cfg = meta["config"]
    snap = product_snapshot_table()[snapshot]
    assert snap["lr"] == cfg["learning_rate"] and snap["num_steps"] == cfg["num_steps"]
    cfg.setdefault("num_segment", 1)                      # (1NoClass has no such attribute: one logit channel)
    assert snap["num_segment"] == cfg["num_segment"]
    params = reference_params(meta)
    layers = [n for n, _ in meta["layers"]]
    keep = [n for n in layers if n not in (meta["seg"], meta["fc"])]
    lr = float(O.poly_lr(cfg["learning_rate"], float(z["in/step"]), cfg["num_steps"]))
    close(lr, z["out/learning_rate"], 1e-6)                # the TF graph computes it in float32
    r = _oracle_step(params, z, meta, snap, lr, keep)
    stats = {n: z["layer_stats"][i] for i, (n, _) in enumerate(meta["layers"])}
    shapes = {n: s for n, s in meta["layers"]}
    compared = 0
    for n in r["__kept__"]:
        assert list(r[n].shape) == shapes[n], (n, r[n].shape, shapes[n])
        close(summary(n, r[n]), stats[n], 1e-9)
        if "layer/" + n in z.files:
            close(r[n], z["layer/" + n], 1e-9)
        compared += 1
    assert compared >= 270, compared                       # every bottleneck convolution / batch norm / junction, the branches, heads
    close(r["seg_logits"], z["out/raw_output_segment"], 1e-9)
    close(r["loss"], z["out/loss"], 1e-11)
    if meta["fc"]:
        close(r["cls_logits"], z["out/raw_output_classes"], 1e-9)
        close(r["loss_segment"], z["out/loss_segment"], 1e-11)
        close(r["loss_classes"], z["out/loss_classes"], 1e-11)
        # the weight of the class term is whatever the reference's add_n used
        w = (float(z["out/loss"]) - float(z["out/loss_segment"])) / float(z["out/loss_classes"])
        assert abs(w - snap["class_weight"]) < 1e-12
    pred, pcls = O.predict_train(r["seg_logits"], r.get("cls_logits"))
    assert np.array_equal(pred, z["out/pred_segment"])
    assert pcls is None or np.array_equal(pcls, z["out/pred_classes"])
    # every gradient and the SGD update
    names = meta["train_op_vars"]
    worst = 0.0
    for i, n in enumerate(names):
        worst = max(worst, close(summary(n, r["grads"][n]), z["grad_stats"][i], 1e-7))
        close(summary(n, r["new_params"][n]), z["new_value_stats"][i], 1e-7)
        if "grad/" + n in z.files:
            close(r["grads"][n], z["grad/" + n], 1e-8)
    print("%s: %d layers, %d gradients; worst gradient-summary error %.2e" % (snapshot, compared, len(names), worst))


def _oracle_step(params, z, meta, snap, lr, keep):
    """oracle.train_step returning every named layer the oracle exposes."""
    cfg = meta["config"]
    exposed = _exposed_layers(params, z, meta)
    have = [n for n in keep if n in exposed]
    r = O.train_step(params, z["in/data"], z["in/label_segment"], z["in/label_classes"], meta["snapshot"],
                     cfg["num_segment"], cfg["last_pool_size"], snap["pos_weight"], snap["class_weight"], lr,
                     torch.float64, cfg.get("attention_class"), keep=tuple(have))
    r["__kept__"] = have
    return r


def _exposed_layers(params, z, meta):
    """Names in the oracle's layer dictionary (it is local to the forward: one run with a spy on the trunk)."""
    cfg = meta["config"]
    rec = {}
    orig = O._pspnet_trunk

    def spy(p, x, L):
        rec["L"] = L
        return orig(p, x, L)
    O._pspnet_trunk = spy
    try:
        O.pspnet_forward(O.to_torch(params, torch.float64), torch.as_tensor(z["in/data"][:1]).to(torch.float64),
                         meta["snapshot"], cfg["num_segment"], cfg["last_pool_size"], cfg.get("attention_class"))
    finally:
        O._pspnet_trunk = orig
    return set(rec["L"].keys())


# --------------------------------------------------------------------------------------------------------------
# product graph builder <-> reference code
# --------------------------------------------------------------------------------------------------------------
def build_product_net(meta):
    from basi_b200.BAISPSPNet import Placeholder, PSPNet
    cfg = meta["config"]
    S = cfg["input_size"][0]
    return PSPNet({"data": Placeholder((None, S, S, 4))}, is_training=True, num_classes=cfg["num_classes"],
                  num_segment=cfg.get("num_segment", 1), last_pool_size=cfg["last_pool_size"],
                  filter_number=cfg["filter_number"], attention_class=cfg.get("attention_class"),
                  variant=meta["snapshot"])


def product_trace(net):
    """The primitive-op sequence of the product's recorded graph, in the vocabulary of the shim's trace."""
    out = []
    for nd in net.nodes:
        a = nd.attrs
        if nd.op == "data":
            continue
        if nd.op == "zero_padding":
            out.append(["pad", {"paddings": [[0, 0], [a["pad"]] * 2, [a["pad"]] * 2, [0, 0]]}])
        elif nd.op == "conv":
            kh, kw, cin, cout = net.variables[a["weights"]]
            out.append(["conv2d", {"k": [kh, kw], "cin": cin, "cout": cout, "strides": [a["stride"]] * 2,
                                   "padding": a["padding"], "rate": a["dilation"]}])
            if a["biases"]:
                out.append(["bias_add", {"c": cout}])
            if a["relu"]:
                out.append(["relu", {}])
        elif nd.op == "batch_normalization":
            out.append(["batch_normalization", {"c": nd.shape[-1], "momentum": a["momentum"], "epsilon": a["epsilon"],
                                                "training": True}])
            if a["relu"]:
                out.append(["relu", {}])
        elif nd.op == "relu":
            out.append(["relu", {}])
        elif nd.op == "max_pool":
            out.append(["max_pool", {"k": [a["k"]] * 2, "strides": [2, 2], "padding": "SAME" if a["k"] == 3 else "VALID"}])
        elif nd.op == "avg_pool":
            out.append(["avg_pool", {"k": [a["k"]] * 2, "strides": [a["k"]] * 2, "padding": "VALID"}])
        elif nd.op == "resize_bilinear":
            out.append(["resize_bilinear", {"align_corners": True}])
        elif nd.op == "concat":
            out.append(["concat", {"axis": -1, "n": len(nd.inputs), "channels": [i.shape[2] for i in nd.inputs]}])
        elif nd.op == "add":
            out.append(["add_n", {"n": len(nd.inputs)}])
        elif nd.op == "multiply":
            out.append(["relu", {}])                       # the ReLU hidden inside Network.multiply
            out.append(["multiply", {"channels": [nd.inputs[0].shape[2], 1]}])
        elif nd.op == "squeeze":
            out.append(["squeeze", {"axis": [1, 2]}])
        elif nd.op == "fc":
            out.append(["relu_layer" if a["relu"] else "xw_plus_b", {"shape": list(net.variables[a["weights"]])}])
        else:
            raise AssertionError("op %s has no reference counterpart in this network" % nd.op)
    return out


@pytest.mark.parametrize("snapshot", SNAPSHOTS)
def test_product_builder_registers_the_reference_codes_graph(snapshot):
    meta, _ = load(snapshot)
    net = build_product_net(meta)
    # layers: same names, same order, same shapes (the explicit zero-padding layers are numbered differently:
    # the reference names them padding1..padding33, the builder after their block)
    ref_layers = [(n, tuple(s[1:])) for n, s in meta["layers"] if not n.startswith("padding")]
    ref_pads = [tuple(s[1:]) for n, s in meta["layers"] if n.startswith("padding")]
    mine = [(n, tuple(nd.shape)) for n, nd in net.layers.items() if n != "data" and not n.startswith("padding")]
    my_pads = [tuple(nd.shape) for n, nd in net.layers.items() if n.startswith("padding")]
    assert mine == ref_layers
    assert my_pads == ref_pads and len(ref_pads) == 33
    # variables: names, shapes, creation order
    ref_vars = [(n, tuple(s)) for n, s, trainable in meta["variables"] if trainable]
    assert list(net.variables.items()) == ref_vars
    # the primitive ops, one for one
    assert product_trace(net) == meta["trace"]
    # ... and what build_net issued after the network: label resize, the two losses with the constants of the
    # product's per-snapshot table, the two-term sum
    snap = product_snapshot_table()[snapshot]
    cfg = meta["config"]
    seg_loss = (["weighted_cross_entropy_with_logits", {"pos_weight": snap["pos_weight"]}] if snap["kind"] == "bce"
                else ["sparse_softmax_cross_entropy_with_logits", {"classes": cfg["num_segment"]}])
    class_part = ([["sparse_softmax_cross_entropy_with_logits", {"classes": cfg["num_classes"]}], ["add_n", {"n": 2}]]
                  if meta["fc"] else [])
    assert meta["trace_train"] == [["resize_nearest_neighbor", {"align_corners": False}], seg_loss] + class_part


@pytest.mark.parametrize("snapshot", SNAPSHOTS)
def test_product_accuracy_fetches_are_the_reference_codes(snapshot):
    """The accuracy_* values Train.run_step reports, computed by the product's host code from the reference run's own
    logits / predictions / labels, equal what the reference's build_net computed."""
    from basi_b200.BAISRunnerTrain import Train
    meta, z = load(snapshot)
    nseg = meta["config"].get("num_segment", 1)
    got = Train.segment_accuracies(nseg, z["out/raw_output_segment"], z["out/pred_segment"], z["in/label_segment"])
    for k, v in got.items():
        assert abs(v - float(z["out/" + k])) < 1e-12, (k, v, float(z["out/" + k]))
    assert set(got) == ({"accuracy_0", "accuracy_1"} if nseg == 1 else {"accuracy_segment"})
    if meta["fc"]:
        assert abs(Train.class_accuracy(z["out/pred_classes"], z["in/label_classes"])
                   - float(z["out/accuracy_classes"])) < 1e-12


@pytest.mark.parametrize("snapshot", SNAPSHOTS)
def test_product_class_only_subset_is_the_reference_var_list(snapshot):
    """A17: `minimize(loss, var_list=[v for v in trainable_variables() if 'class_attention' in v.name])`."""
    meta, _ = load(snapshot)
    net = build_product_net(meta)
    # Engine.set_trainable('class_attention') selects by the same substring rule over the same names
    assert [n for n in net.variables if "class_attention" in n] == meta["train_classes_op_vars"]
    assert len(meta["train_classes_op_vars"]) == (4 if meta["fc"] else 0)


# --------------------------------------------------------------------------------------------------------------
# F4: the cascade of back/8AttentionU -- the reference's whole Train.__init__ (its own Data reader on voc_mini,
# BAISNet(...).build(), cal_loss, both optimizers) at its own filter_number = 32
# --------------------------------------------------------------------------------------------------------------
def test_cascade_inventory_is_the_reference_codes():
    meta, _ = load("8AttentionU")
    cfg = meta["config"]
    ref = [(n, tuple(s)) for n, s, trainable in meta["variables"] if trainable]
    specs = O.attention_u_specs(cfg["num_classes"], cfg["num_segment"], cfg["filter_number"],
                                cfg["attention_module_num"])
    assert [(n, tuple(s)) for n, s in specs.items()] == ref
    assert meta["train_op_vars"] == [n for n, _ in ref]
    # "单独训练最后的attention": var_list = names containing 'attention' (BAISRunnerTrain.py:91-95)
    assert meta["train_attention_op_vars"] == [n for n, _ in ref if "attention" in n]
    from basi_b200.BAISNet import BAISNet
    from basi_b200.BAISPSPNet import Placeholder
    S = cfg["input_size"][0]
    net = BAISNet(Placeholder((None, S, S, 4)), is_training=True, num_classes=cfg["num_classes"],
                  num_segment=cfg["num_segment"], segment_attention=cfg["segment_attention"],
                  last_pool_size=cfg["last_pool_size"], filter_number=cfg["filter_number"],
                  attention_module_num=cfg["attention_module_num"])
    assert list(net.variables.items()) == ref
    # the arithmetic ops, one for one (the reference additionally builds an unused ReLU of every decoder output and
    # takes softmax -> split -> multiply where the builder records softmax_gate -> mask_multiply: compared by value
    # in the oracle test below and in tests/test_gpu_net.py)
    skip = ("relu", "softmax", "multiply")
    k = [t[0] for t in meta["trace"]].index("resize_nearest_neighbor")      # the network ends, the loss code begins
    ref_ops = [t for t in meta["trace"][:k] if t[0] not in skip]
    mine = []
    for nd in net.nodes:
        if nd.op in ("softmax_gate", "mask_multiply"):
            continue
        if nd.op == "sigmoid":
            mine.append(["sigmoid", {}])
            continue
        sub = type("N", (), {"nodes": [nd], "variables": net.variables})()
        mine += [t for t in product_trace(sub) if t[0] not in skip]
    assert mine == ref_ops
    nc = cfg["num_classes"]
    assert meta["trace"][k:k + 12] == (
        [["resize_nearest_neighbor", {"align_corners": False}]] * 2
        + [["sparse_softmax_cross_entropy_with_logits", {"classes": cfg["num_segment"]}]] * 2
        + [["weighted_cross_entropy_with_logits", {"pos_weight": 3.0}]] * 2
        + [["sparse_softmax_cross_entropy_with_logits", {"classes": nc}]] * 4 + [["add_n", {"n": 4}]] * 2)


def test_cascade_oracle_train_step_reproduces_the_reference_code():
    meta, z = load("8AttentionU")
    cfg = meta["config"]
    snap = product_snapshot_table()["8AttentionU"]
    assert snap["lr"] == cfg["learning_rate"] and snap["num_steps"] == cfg["num_steps"]
    params = reference_params(meta)
    lr = float(O.poly_lr(cfg["learning_rate"], float(z["in/step"]), cfg["num_steps"]))
    close(lr, z["out/learning_rate"], 1e-6)
    # the batch came out of the reference's own reader: attention labels are (label == 1), labels are 0..3
    assert np.array_equal(z["in/label_attention"], (z["in/label_segment"] == 1).astype(np.int64))
    assert z["in/label_segment"].max() <= 3 and z["in/data"].dtype == np.float32
    r = O.attention_u_train_step(params, z["in/data"], z["in/label_segment"], z["in/label_attention"],
                                 z["in/label_classes"], cfg["last_pool_size"], lr, torch.float64, cfg["num_segment"],
                                 cfg["segment_attention"], cfg["attention_module_num"])
    for i in range(4):
        close(r["segments"][i], z["out/segment_%d" % i], 1e-9)
        close(np.transpose(r["attentions"][i], (0, 2, 3, 1)), z["out/attention_%d" % i], 1e-9)
        close(r["classes"][i], z["out/class_%d" % i], 1e-9)
    close(r["loss"], z["out/loss"], 1e-11)
    close(r["loss_segment"], z["out/loss_segment_all"], 1e-11)
    close(r["loss_classes"], z["out/loss_class_all"], 1e-11)
    w = (float(z["out/loss"]) - float(z["out/loss_segment_all"])) / float(z["out/loss_class_all"])
    assert abs(w - snap["class_weight"]) < 1e-12
    assert np.array_equal(np.argmax(r["segments"][0], -1)[..., None], z["out/pred_segment"])
    assert np.array_equal(np.argmax(r["classes"][0], -1), z["out/pred_classes"])
    worst = 0.0
    for i, n in enumerate(meta["train_op_vars"]):
        worst = max(worst, close(summary(n, r["grads"][n]), z["grad_stats"][i], 1e-7))
        close(summary(n, r["new_params"][n]), z["new_value_stats"][i], 1e-7)
        if "grad/" + n in z.files:
            close(r["grads"][n], z["grad/" + n], 1e-8)
    print("8AttentionU: %d gradients; worst gradient-summary error %.2e" % (len(meta["train_op_vars"]), worst))


# --------------------------------------------------------------------------------------------------------------
# F2: the current-HEAD path -- the reference's whole Train.__init__ (BAISRunnerTrain.py + BAISNet.py + BAISData.py +
# slim/nets/nets_factory.py -> slim/nets/vgg.py) at 224^2 on its own reader's batch of tests/golden/voc_mini
# --------------------------------------------------------------------------------------------------------------
def _head_inputs(z):
    image = z["in/image_u8"].astype(np.float32) / np.float32(255)          # Data._read_image: uint8 / 255 in float32
    return image, z["in/label_segment"].astype(np.int64)


def test_head_inventory_is_the_reference_codes():
    meta, z = load("HEAD")
    trained = [(n, tuple(s)) for n, s, trainable in meta["variables"] if trainable and n in set(meta["train_op_vars"])]
    specs = O.linknet_top_specs(1.0)
    assert [(n, tuple(s)) for n, s in specs.items()] == trained
    # the rest of what slim's vgg_16 builds (fc6 as a 7x7 VALID convolution, fc7) is in the reference's graph but the
    # loss does not depend on it: minimize() has no gradient for it, no fetched value reads it
    rest = [n for n, _, t in meta["variables"] if n not in set(meta["train_op_vars"])]
    assert rest == ["vgg_16/fc6/weights", "vgg_16/fc6/biases", "vgg_16/fc7/weights", "vgg_16/fc7/biases"]
    assert meta["dropout_calls"] == 1                                  # ... nor on vgg_16's dropout6
    assert meta["train_segment_side_op_vars"] == [n for n, _ in trained if "segment_side" in n]
    from basi_b200.BAISNet import LinkNetTop
    from basi_b200.BAISPSPNet import Placeholder
    S = meta["config"]["input_size"][0]
    net = LinkNetTop(Placeholder((None, S, S, 3)))
    assert list(net.variables.items()) == trained
    segs, _ = net.build()
    assert [[meta["config"]["batch_size"]] + list(nd.shape) for nd in segs] == meta["segment_shapes"]
    # primitive ops of the path: the reference's trace minus pool5 / fc6 / dropout6 / fc7 and minus the loss code
    ops = meta["trace"]
    k = [t[0] for t in ops].index("dropout")
    dead = ops[k - 4:k + 4]                                             # pool5, fc6 (+bias +relu), dropout, fc7 (+bias +relu)
    assert [t[0] for t in dead] == ["max_pool", "conv2d", "bias_add", "relu", "dropout", "conv2d", "bias_add", "relu"]
    live = ops[:k - 4] + ops[k + 4:]
    n_loss = [t[0] for t in live].index("weighted_cross_entropy_with_logits") - 1
    mine = []
    for nd in net.nodes:
        if nd.op == "resize_nearest":
            mine.append(["resize_nearest_neighbor", {"align_corners": False}])
            continue
        sub = type("N", (), {"nodes": [nd], "variables": net.variables})()
        mine += product_trace(sub)
    assert mine == live[:n_loss]
    assert [t for t in live[n_loss:] if t[0] != "resize_nearest_neighbor"][:6] == (
        [["weighted_cross_entropy_with_logits", {"pos_weight": 1.0}]] * 5 + [["add_n", {"n": 5}]])


def test_head_oracle_train_step_reproduces_the_reference_code():
    meta, z = load("HEAD")
    cfg = meta["config"]
    params = {n: param_value(n, s, kind_of(n)) for n, s, _ in meta["variables"] if n in set(meta["train_op_vars"])}
    image, label = _head_inputs(z)
    # HEAD's schedule differs from the snapshots': power 0.8, 100001 steps (BAISRunnerTrain.py:33,49-50)
    lr = float(O.poly_lr(cfg["learning_rate"], float(z["in/step"]), cfg["num_steps"], power=0.8))
    close(lr, z["out/learning_rate"], 1e-6)
    from basi_b200.BAISRunnerTrain import HEAD_SCHEDULE, poly_learning_rate
    close(poly_learning_rate(step=float(z["in/step"]), **HEAD_SCHEDULE), z["out/learning_rate"], 1e-6)
    r = O.linknet_top_train_step(params, image, label, lr, torch.float64, pos_weight=1.0)
    assert [list(a.shape) for a in r["segments"]] == meta["segment_shapes"]
    for i, a in enumerate(r["segments"]):
        close(summary("segment_%d" % i, a), z["out/segment_stats_%d" % i], 1e-9)
        if "out/segment_%d" % i in z.files:
            close(a, z["out/segment_%d" % i], 1e-9)
        close(r["loss_segments"][i], z["out/loss_segment_%d" % i], 1e-11)
    close(r["loss"], z["out/loss"], 1e-11)
    worst = 0.0
    for i, n in enumerate(meta["train_op_vars"]):
        worst = max(worst, close(summary(n, r["grads"][n]), z["grad_stats"][i], 1e-7))
        close(summary(n, r["new_params"][n]), z["new_value_stats"][i], 1e-7)
        if "grad/" + n in z.files:
            close(r["grads"][n], z["grad/" + n], 1e-8)
    print("HEAD: %d gradients; worst gradient-summary error %.2e" % (len(meta["train_op_vars"]), worst))


# --------------------------------------------------------------------------------------------------------------
# F1: variant B (back/90AttentionSingle2) -- the reference's whole Train.__init__ at the 720^2 its class head is
# hard-coded for (p_size=15, k_size=3 on the 45x45 block4), batch 1 from its own reader on tests/golden/voc_mini
# --------------------------------------------------------------------------------------------------------------
def _variant_b_inputs(z):
    image = z["in/image_u8"].astype(np.float32) / np.float32(255)
    return image, z["in/mask"], z["in/label_segment"].astype(np.int64), z["in/label_classes"]


def test_variant_b_inventory_is_the_reference_codes():
    meta, z = load("90AttentionSingle2")
    with_grad = set(meta["train_op_vars"])
    trainable = [(n, tuple(s)) for n, s, t in meta["variables"] if t]
    specs = O.linknet_b_specs(meta["config"]["num_classes"], 1.0)
    on_path = [(n, s) for n, s in trainable if not n.startswith(("vgg_16/fc6", "vgg_16/fc7"))]
    assert [(n, tuple(s)) for n, s in specs.items()] == on_path
    # no gradient: vgg_16's fc6 / fc7 (never read) and the finest attention block's a_conv_o (its output feeds nothing)
    assert [n for n, _ in trainable if n not in with_grad] == [
        "vgg_16/fc6/weights", "vgg_16/fc6/biases", "vgg_16/fc7/weights", "vgg_16/fc7/biases",
        "attention_1/attention_1_attention/a_conv_o/weights"]
    assert meta["segments"] == 0 and meta["classes"] == 1
    assert meta["train_attention_op_vars"] == [n for n, _ in trainable if "attention" in n and n in with_grad]
    from basi_b200.BAISNet import LinkNet
    from basi_b200.BAISPSPNet import Placeholder
    S = meta["config"]["input_size"][0]
    net = LinkNet(Placeholder((None, S, S, 3)), Placeholder((None, S, S, 1), name="mask"), True,
                  num_classes=meta["config"]["num_classes"], p_size=15, k_size=3)
    assert list(net.variables.items()) == on_path
    _, atts, clss = net.build()
    assert [[1] + list(nd.shape) for nd in atts] == meta["attention_shapes"] and len(clss) == 1
    snap = product_snapshot_table()["90AttentionSingle2"]
    assert snap["lr"] == meta["config"]["learning_rate"] and snap["num_steps"] == meta["config"]["num_steps"]


def test_variant_b_oracle_train_step_reproduces_the_reference_code():
    meta, z = load("90AttentionSingle2")
    cfg = meta["config"]
    params = {n: param_value(n, s, kind_of(n)) for n, s, t in meta["variables"]
              if t and not n.startswith(("vgg_16/fc6", "vgg_16/fc7"))}
    image, mask, label, cls = _variant_b_inputs(z)
    lr = float(O.poly_lr(cfg["learning_rate"], float(z["in/step"]), cfg["num_steps"]))
    close(lr, z["out/learning_rate"], 1e-6)
    r = O.linknet_b_train_step(params, image, mask, label, cls, lr, torch.float64, p_size=15, pos_weight=3.0)
    assert [list(a.shape) for a in r["attentions"]] == meta["attention_shapes"]
    for i, a in enumerate(r["attentions"]):
        close(summary("attention_%d" % i, a), z["out/attention_stats_%d" % i], 1e-9)
        if "out/attention_%d" % i in z.files:
            close(a, z["out/attention_%d" % i], 1e-9)
    close(r["cls_logits"], z["out/class_0"], 1e-9)
    close(r["loss"], z["out/loss"], 1e-11)
    close(r["loss_attention"], z["out/loss_segment_all"], 1e-11)
    close(r["loss_classes"], z["out/loss_class_all"], 1e-11)
    worst = 0.0
    for i, n in enumerate(meta["train_op_vars"]):
        worst = max(worst, close(summary(n, r["grads"][n]), z["grad_stats"][i], 1e-7))
        close(summary(n, r["new_params"][n]), z["new_value_stats"][i], 1e-7)
        if "grad/" + n in z.files:
            close(r["grads"][n], z["grad/" + n], 1e-8)
    # the variable without a gradient stays where it was
    n0 = "attention_1/attention_1_attention/a_conv_o/weights"
    assert not np.any(r["grads"][n0]) and np.array_equal(r["new_params"][n0], params[n0].astype(np.float64))
    print("90AttentionSingle2: %d gradients; worst gradient-summary error %.2e" % (len(meta["train_op_vars"]), worst))


# --------------------------------------------------------------------------------------------------------------
# the readers of variant B and of the current-HEAD script: the batches the reference's own readers fed to the runs above
# --------------------------------------------------------------------------------------------------------------
VOC_MINI = os.path.join(HERE, "golden", "voc_mini") + "/"


def test_variant_b_reader_reproduces_the_reference_readers_batch():
    """back/90AttentionSingle2/BAISData.py on tests/golden/voc_mini at 720^2: image, full-resolution click map (bit for
    bit), {0,1} attention labels with the border ring as background, class id -- the batch of the variant-B fixture."""
    from basi_b200.BAISData import DataAttention
    meta, z = load("90AttentionSingle2")
    S = meta["config"]["input_size"]
    reader = DataAttention(data_root_path=VOC_MINI, data_list="ImageSets/Segmentation/train.txt", batch_size=1,
                           image_size=S)
    np.random.seed(9)                                      # the seed the generator set before next_batch_train()
    data, mask, att, cls = reader.next_batch_train()
    assert np.array_equal(np.asarray(data, np.float32), z["in/image_u8"].astype(np.float32) / np.float32(255))
    assert np.array_equal(np.asarray(mask, np.float32).view(np.uint32), z["in/mask"].view(np.uint32))
    assert np.asarray(att).dtype == np.int32 and np.array_equal(np.asarray(att), z["in/label_segment"])
    assert [int(c) for c in cls] == [int(c) for c in z["in/label_classes"]]
    assert reader.number_patch == len(reader._annotations) and len(reader._annotations) == 8      # instances, not images


def test_head_reader_reproduces_the_reference_readers_batch():
    """BAISData.py (HEAD) on tests/golden/voc_mini at 224^2: one foreground map per image, the 255 border ring counted
    as foreground -- the batch of the HEAD fixture."""
    from basi_b200.BAISData import DataTop
    meta, z = load("HEAD")
    reader = DataTop(data_root_path=VOC_MINI, data_list="ImageSets/Segmentation/train.txt",
                     batch_size=meta["config"]["batch_size"], image_size=meta["config"]["input_size"])
    data, ann = reader.next_batch_train()
    assert np.array_equal(np.asarray(data, np.float32), z["in/image_u8"].astype(np.float32) / np.float32(255))
    assert np.array_equal(np.asarray(ann), z["in/label_segment"]) and set(np.unique(np.asarray(ann))) == {0, 1}
    assert reader.number_patch == 1                        # 3 images, batch 2


def test_head_training_wrapper_builds_and_feeds_like_the_reference_script(tmp_path):
    """TrainTop (the call surface of the HEAD script's Train) in dry-run mode: the plan is the top-level LinkNet's, the
    reader's batch fits the engine's input / full-resolution label buffers, the learning rate is the one the reference's
    script computed, the segment_side subset is the script's var_list."""
    from basi_b200.BAISRunnerTrain import TrainTop
    meta, z = load("HEAD")
    S, B = meta["config"]["input_size"], meta["config"]["batch_size"]
    tr = TrainTop(batch_size=B, input_size=S, log_dir=str(tmp_path / "log"), data_root_path=VOC_MINI,
                  train_list="ImageSets/Segmentation/train.txt", data_path="JPEGImages/",
                  annotation_path="SegmentationObject/", class_path="SegmentationClass/", precision="f32", dry_run=True)
    eng = tr.engine
    assert tuple(eng.input.t.shape) == (B, S[0], S[1], 3) and tuple(eng.label_seg.shape) == (B, S[0], S[1], 1)
    assert list(eng.param_index.keys()) == meta["train_op_vars"]
    lr = tr.feed(float(z["in/step"]))
    close(lr, z["out/learning_rate"], 1e-6)
    assert np.array_equal(eng.input.t.numpy(), z["in/image_u8"].astype(np.float32) / np.float32(255))
    assert np.array_equal(eng.label_seg.numpy(), z["in/label_segment"].astype(np.float32))
    assert eng.set_trainable("segment_side") >= 1
    assert [n for n in eng.param_index if "segment_side" in n] == meta["train_segment_side_op_vars"]
    assert (tr.learning_rate, tr.num_steps, tr.cal_step) == (5e-3, 100001, 1)
    assert [[B] + list(nd.shape) for nd in tr.segments] == meta["segment_shapes"]


# --------------------------------------------------------------------------------------------------------------
# cfg1, the reference's own CPU-runnable case: back/4BorderClass/BAISRunnerOne.py Runner.run on input/7.jpg
# --------------------------------------------------------------------------------------------------------------
def _runner_one_case():
    meta, z = load("RunnerOne")
    from basi_b200.BAISData import Data
    cfg = meta["config"]
    where = [int(v) for v in z["in/where"]]
    loaded = Data.load_image(os.path.join(HERE, "golden", "input_7.jpg"), where=where, image_size=cfg["input_size"])
    data = np.asarray(loaded[0], dtype=np.float32)
    params = {n: param_value(n, s, kind_of(n)) for n, s, t in meta["variables"] if t}
    return meta, z, cfg, where, data, params


def test_runner_one_oracle_reproduces_the_reference_script():
    """Click inference of the reference's own script on its own fixture image: click map (bit for bit), 4-channel
    logits, sigmoid -> argmax mask, class logits / argmax, and the PNG files the script wrote."""
    meta, z, cfg, where, data, params = _runner_one_case()
    assert np.array_equal(data[0, :, :, 3].view(np.uint32), z["in/click_map"].view(np.uint32))
    assert np.array_equal(O.mask_gaussian(cfg["input_size"], where).view(np.uint32), z["in/click_map"].view(np.uint32))
    with torch.no_grad():
        out = O.pspnet_forward(O.to_torch(params, torch.float64), torch.as_tensor(data).to(torch.float64),
                               "4BorderClass", cfg["num_segment"], cfg["last_pool_size"])
    logits = out["conv6_n_4"].numpy()
    close(logits, z["out/raw_output"], 1e-9)
    close(out["class_attention_fc"].numpy(), z["out/raw_output_classes"], 1e-9)
    assert np.array_equal(O.predict_click(z["out/raw_output"]), z["out/predict_output"])
    assert np.array_equal(np.argmax(logits, -1), z["out/predict_output"])
    assert int(np.argmax(out["class_attention_fc"].numpy(), -1)[0]) == int(z["out/pred_classes"][0])
    # the files the reference script wrote: <name>pred.png = argmax * 255 // 4, <name>pred_k.png = uint8(sigmoid_k * 255)
    name = [f for f in meta["files"] if f.endswith("pred.png")][0]
    assert np.array_equal(z["file/" + name], (z["out/predict_output"][0] * 255 // 4).astype(np.uint8))
    sig = 1.0 / (1.0 + np.exp(-logits[0]))
    for k in range(4):
        ref = z["file/" + name.replace("pred.png", "pred_%d.png" % k)]
        mine = np.asarray(sig[:, :, k] * 255, dtype=np.uint8)
        assert np.mean(mine != ref) < 1e-3 and np.abs(mine.astype(int) - ref.astype(int)).max() <= 1


@pytest.mark.gpu
def test_cuda_runner_one_matches_the_reference_script(tmp_path):
    """The product's Runner.run (f32 mode) on tests/golden/input_7.jpg with the reference run's parameters restored from
    a TF-named checkpoint: logits within 1e-4 of what the reference's own script computed, the same class, the mask
    and the written pred.png equal except where two logits tie within float32."""
    from PIL import Image
    from basi_b200.BAISRunnerOne import Runner
    meta, z, cfg, where, data, params = _runner_one_case()
    log_dir, save_dir = str(tmp_path / "model"), str(tmp_path / "out")
    os.makedirs(log_dir)
    np.savez(os.path.join(log_dir, "model.ckpt-0.npz"), **params)
    res = Runner(log_dir=log_dir, save_dir=save_dir, precision="f32").run(
        result_filename="7_360_480_", image_filename=os.path.join(HERE, "golden", "input_7.jpg"), where=where)
    e = _rel(res["raw_output"], z["out/raw_output"])
    assert e < F32_TOL, e
    assert int(res["pred_classes"][0]) == int(z["out/pred_classes"][0])
    assert _rel(res["raw_output_classes"], z["out/raw_output_classes"]) < 10 * F32_TOL
    agree = float(np.mean(res["predict_output"] == z["out/predict_output"]))
    assert agree >= 0.999, agree
    name = [f for f in meta["files"] if f.endswith("pred.png")][0]
    png = np.asarray(Image.open(os.path.join(save_dir, name)))
    assert float(np.mean(png == z["file/" + name])) >= 0.999
    print("RunnerOne: CUDA f32 vs the reference script: logits %.2e, mask agreement %.5f" % (e, agree))


# --------------------------------------------------------------------------------------------------------------
# the one-logit snapshots' inference script: back/2AddClass/BAISRunnerOne.py at its own 400^2, click sampled from an
# instance annotation of tests/golden/voc_mini
# --------------------------------------------------------------------------------------------------------------
def _runner_one_2addclass_case():
    meta, z = load("RunnerOne_2AddClass")
    cfg = meta["config"]
    params = {n: param_value(n, s, kind_of(n)) for n, s, t in meta["variables"] if t}
    image = os.path.join(HERE, "golden", cfg["image"])
    ann = os.path.join(HERE, "golden", cfg["annotation"])
    return meta, z, cfg, params, image, ann


def test_runner_one_2addclass_oracle_reproduces_the_reference_script():
    from basi_b200.BAISData import Data
    meta, z, cfg, params, image, ann = _runner_one_2addclass_case()
    np.random.seed(cfg["numpy_seed"])                      # the script draws the click with np.random.randint
    loaded = Data.load_image(image, annotation_filename=ann, ann_index=cfg["ann_index"], image_size=cfg["input_size"])
    data = np.asarray(loaded[0], dtype=np.float32)
    assert np.array_equal(data[0, :, :, 3].view(np.uint32), z["in/click_map"].view(np.uint32))      # same click, same map
    assert np.array_equal(np.asarray(np.squeeze(loaded[4] * 255), dtype=np.uint8), z["file/b_ann.bmp"])
    with torch.no_grad():
        out = O.pspnet_forward(O.to_torch(params, torch.float64), torch.as_tensor(data).to(torch.float64),
                               "2AddClass", 1, cfg["last_pool_size"])
    raw = out["conv6_n"].numpy()
    close(raw, z["out/raw_output"], 1e-9)
    close(out["class_attention_fc"].numpy(), z["out/raw_output_classes"], 1e-9)
    # training-time rule (logit > 0.5) and the runner's rule (sigmoid > 0.5 <=> logit > 0) are different masks
    pred, _ = O.predict_train(z["out/raw_output"])
    assert np.array_equal((pred[0, :, :, 0] * 255).astype(np.uint8), z["file/b_pred_raw.png"])
    assert np.array_equal(((z["out/raw_output"][0, :, :, 0] > 0) * 255).astype(np.uint8), z["file/b_pred_sigmoid.png"])
    assert (z["file/b_pred_raw.png"] != z["file/b_pred_sigmoid.png"]).sum() > 100


@pytest.mark.gpu
def test_cuda_runner_one_2addclass_matches_the_reference_script(tmp_path):
    """The product's Runner.run(variant='2AddClass', last_pool_size=50) in f32 mode against the files the reference
    script wrote: pred.png (sigmoid grey levels), pred_raw.png (logit > 0.5), pred_sigmoid.png, mask.bmp, ann.bmp."""
    from PIL import Image
    from basi_b200.BAISRunnerOne import Runner
    meta, z, cfg, params, image, ann = _runner_one_2addclass_case()
    log_dir, save_dir = str(tmp_path / "model"), str(tmp_path / "out")
    os.makedirs(log_dir)
    np.savez(os.path.join(log_dir, "model.ckpt-0.npz"), **params)
    np.random.seed(cfg["numpy_seed"])
    res = Runner(log_dir=log_dir, save_dir=save_dir, last_pool_size=cfg["last_pool_size"], variant="2AddClass",
                 num_segment=1, precision="f32").run(result_filename="b_", image_filename=image,
                                                     annotation_filename=ann, ann_index=cfg["ann_index"])
    e = _rel(res["raw_output"].reshape(z["out/raw_output"].shape), z["out/raw_output"])
    assert e < F32_TOL, e
    assert int(res["pred_classes"][0]) == int(z["out/pred_classes"][0])
    worst = 1.0
    for f in meta["files"]:
        if f == "b_data.png":
            continue
        mine = np.asarray(Image.open(os.path.join(save_dir, f)))
        same = float(np.mean(np.abs(mine.astype(int) - z["file/" + f].astype(int)) <= (1 if f == "b_pred.png" else 0)))
        assert same >= 0.999, (f, same)
        worst = min(worst, same)
    print("RunnerOne 2AddClass: CUDA f32 vs the reference script: logits %.2e, files agree on >= %.5f of the pixels"
          % (e, worst))


# --------------------------------------------------------------------------------------------------------------
# cfg1 as the interactive tool runs it: back/4BorderClass/BAISRunnerGUI.py RunnerGUI.run, one click on input/7.jpg
# --------------------------------------------------------------------------------------------------------------
def _runner_gui_case():
    from PIL import Image
    meta, z = load("RunnerGUI")
    image_data = np.array(Image.open(os.path.join(HERE, "golden", "input_7.jpg")))
    params = {n: param_value(n, s, kind_of(n)) for n, s, t in meta["variables"] if t}
    return meta, z, image_data, params


def test_runner_gui_oracle_and_host_post_processing_reproduce_the_reference_tool():
    """What the reference's GUI tool computed and DREW for one click: click position, upsampled argmax(sigmoid) map,
    class, the mask at the displayed size and the blended image handed to plt.imshow."""
    from basi_b200.BAISData import CategoryNames, Data
    from basi_b200.BAISRunnerOne import RunnerGUI
    meta, z, image_data, params = _runner_gui_case()
    cfg = meta["config"]
    point = [int(v) for v in z["in/point_xy"]]
    where = RunnerGUI.click_position(cfg["input_size"], image_data, point)
    assert where == [int(v) for v in z["in/where"]]
    data = np.asarray(Data.load_image(image_data, where=where, image_size=cfg["input_size"])[0], dtype=np.float32)
    with torch.no_grad():
        out = O.pspnet_forward(O.to_torch(params, torch.float64), torch.as_tensor(data).to(torch.float64),
                               "4BorderClass", cfg["num_segment"], cfg["last_pool_size"])
    pred = O.predict_click(out["conv6_n_4"].numpy(), tuple(cfg["input_size"]))
    assert np.mean(pred == z["out/predict_output"]) >= 0.99999          # (legacy bilinear in float32 vs float64: ties)
    cls = int(np.argmax(out["class_attention_fc"].numpy(), -1)[0])
    assert cls == int(z["out/pred_classes"][0]) and CategoryNames[cls] == meta["class_name_drawn"]
    # host post-processing of the product on the reference's own argmax map: exactly the image the tool displayed
    segment = np.squeeze(np.asarray(np.where(z["out/predict_output"][0] == 1, 1, 0), dtype=np.uint8))
    small = RunnerGUI.mask_to_image_size(segment, image_data)
    assert np.array_equal(RunnerGUI.blend(image_data, small, cfg["mask_color"], cfg["opacity"]), z["out/blended"])


@pytest.mark.gpu
def test_cuda_runner_gui_matches_the_reference_tool():
    """The product's RunnerGUI.run_image (f32 mode, CUDA graph, device-side click map and legacy-bilinear argmax) on
    tests/golden/input_7.jpg against what the reference's tool computed and drew for the same click."""
    from basi_b200.BAISRunnerOne import RunnerGUI
    meta, z, image_data, params = _runner_gui_case()
    cfg = meta["config"]
    gui = RunnerGUI(None, last_pool_size=cfg["last_pool_size"], variant="4BorderClass", num_classes=cfg["num_classes"],
                    num_segment=cfg["num_segment"], filter_number=cfg["filter_number"], precision="f32")
    gui.engine.set_params(params)
    seg, cls, where = gui.run_image(os.path.join(HERE, "golden", "input_7.jpg"), [int(v) for v in z["in/point_xy"]])
    assert where == [int(v) for v in z["in/where"]] and cls == int(z["out/pred_classes"][0])
    full = np.asarray(np.where(gui.mask_dev.cpu().numpy()[0] == 1, 1, 0), dtype=np.uint8)
    ref_full = np.squeeze(np.asarray(np.where(z["out/predict_output"][0] == 1, 1, 0), dtype=np.uint8))
    agree = float(np.mean(full == ref_full))
    assert agree >= 0.999, agree
    blended = RunnerGUI.blend(image_data, seg, cfg["mask_color"], cfg["opacity"])
    same = float(np.mean(np.all(blended == z["out/blended"], axis=-1)))
    assert same >= 0.999, same
    print("RunnerGUI: CUDA f32 vs the reference tool: mask agreement %.5f at 720^2, displayed image %.5f" % (agree, same))


# --------------------------------------------------------------------------------------------------------------
# F2 inference: the reference's BAISRunnerTest.Inference(...).load_model() / .inference(...) on input/7.jpg
# --------------------------------------------------------------------------------------------------------------
def _head_inference_case():
    meta, z = load("HEAD_Inference")
    params = {n: param_value(n, s, kind_of(n)) for n, s, t in meta["variables"]
              if t and not n.startswith(("vgg_16/fc6", "vgg_16/fc7"))}
    for k in z.files:
        if k.startswith("param_override/"):
            params[k[len("param_override/"):]] = z[k]
    return meta, z, params


def test_head_inference_oracle_reproduces_the_reference_script():
    from basi_b200.BAISRunnerTest import Inference
    meta, z, params = _head_inference_case()
    S = meta["config"]["input_size"]
    im = Inference.load_data(os.path.join(HERE, "golden", "input_7.jpg"), S)
    with torch.no_grad():
        segs = O.linknet_top_forward(O.to_torch(params, torch.float64), torch.as_tensor(im[None]).to(torch.float64))
    close(segs[0].numpy(), z["out/segment_0"], 1e-9)                   # => load_data decoded the same pixels, too
    pred = np.argmax(segs[0].numpy(), -1).astype(np.uint8)
    assert np.array_equal(pred, z["out/pred_segment"]) and 0.2 < pred.mean() < 0.8
    assert np.array_equal(z["file/7.bmp"], pred[0] * 255)              # the .bmp the script saved


@pytest.mark.gpu
def test_cuda_head_inference_matches_the_reference_script(tmp_path):
    """The product's BAISRunnerTest.Inference (f32 mode) with the reference run's parameters restored from a TF-named
    checkpoint: the mask of the coarsest head and the saved 7.bmp against the reference script's."""
    from PIL import Image
    from basi_b200.BAISRunnerTest import Inference
    meta, z, params = _head_inference_case()
    log_dir = str(tmp_path / "model")
    os.makedirs(log_dir)
    np.savez(os.path.join(log_dir, "model.ckpt-0.npz"), **params)
    inf = Inference(meta["config"]["input_size"], str(tmp_path / "summary"), log_dir, precision="f32")
    assert inf.load_model() is not None
    pred = inf.inference(os.path.join(HERE, "golden", "input_7.jpg"), 0, save_path=str(tmp_path / "out"))
    logits = inf.engine.att_logits[0].t.float().cpu().numpy()
    e = _rel(logits.reshape(z["out/segment_0"].shape), z["out/segment_0"])
    assert e < F32_TOL, e
    agree = float(np.mean(pred == z["out/pred_segment"][0]))
    assert agree >= 0.995, agree                                       # 676 pixels: at most 3 float32 ties
    bmp = np.asarray(Image.open(os.path.join(str(tmp_path / "out"), "input_7.bmp")))    # (named after the image file)
    assert float(np.mean(bmp == z["file/7.bmp"])) >= 0.995
    print("HEAD Inference: CUDA f32 vs the reference script: logits %.2e, mask agreement %.4f" % (e, agree))


# --------------------------------------------------------------------------------------------------------------
# CUDA f32 path <-> reference code (runs last in the -m gpu suite)
# --------------------------------------------------------------------------------------------------------------
F32_TOL = 1e-4          # BASELINE.json north_star: float32 within 1e-4 relative


def _rel(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return float(np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-30))


def _rel2(a, b):
    a, b = np.asarray(a, np.float64).reshape(-1), np.asarray(b, np.float64).reshape(-1)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30))


@pytest.mark.gpu
@pytest.mark.parametrize("snapshot", SNAPSHOTS)
def test_cuda_train_step_matches_the_reference_code(snapshot):
    """The f32 mode of the CUDA path (tcgen05 split operands) through the C ABI against what the reference's own
    build_net computed: layers, logits, predictions, losses, gradients, SGD update."""
    from basi_b200.BAISPSPNet import Placeholder, PSPNet           # noqa: F401
    from basi_b200.engine import Engine
    meta, z = load(snapshot)
    cfg = meta["config"]
    cfg.setdefault("num_segment", 1)
    snap = product_snapshot_table()[snapshot]
    B = cfg["batch_size"]
    params = reference_params(meta)
    net = build_product_net(meta)
    eng = Engine(net, B, "f32", True, dict(kind=snap["kind"], pos_weight=snap["pos_weight"],
                                           class_weight=snap["class_weight"]))
    eng.set_params(params)
    from basi_b200.BAISRunnerTrain import poly_learning_rate
    lr = float(poly_learning_rate(snap["lr"], float(z["in/step"]), snap["num_steps"]))
    assert abs(lr - float(z["out/learning_rate"])) < 1e-6 * lr
    eng.feed(z["in/data"], z["in/label_segment"], z["in/label_classes"], lr)
    eng.step_device()
    torch.cuda.synchronize()
    # float32 noise floor of this input: the oracle in float32 against the reference-code numbers
    r32 = O.train_step(params, z["in/data"], z["in/label_segment"], z["in/label_classes"], snapshot,
                       cfg["num_segment"], cfg["last_pool_size"], snap["pos_weight"], snap["class_weight"], lr,
                       torch.float32, cfg.get("attention_class"))
    loss, lseg, lcls = eng.losses()
    assert abs(loss - float(z["out/loss"])) < F32_TOL * max(1, abs(float(z["out/loss"])))
    logits = eng.seg_logits.t.cpu().numpy()
    e_seg = _rel(logits.reshape(z["out/raw_output_segment"].shape), z["out/raw_output_segment"])
    assert e_seg < F32_TOL, e_seg
    assert np.array_equal(eng.pred_seg.cpu().numpy().reshape(-1), z["out/pred_segment"].reshape(-1))
    e_cls = 0.0
    if meta["fc"]:
        assert abs(lseg - float(z["out/loss_segment"])) < F32_TOL * max(1, abs(float(z["out/loss_segment"])))
        assert abs(lcls - float(z["out/loss_classes"])) < F32_TOL * max(1, abs(float(z["out/loss_classes"])))
        cl = eng.cls_logits.t.cpu().numpy().reshape(B, -1)
        e_cls = _rel(cl, z["out/raw_output_classes"])
        assert e_cls < F32_TOL + 3 * _rel(r32["cls_logits"], z["out/raw_output_classes"]), e_cls
        assert np.array_equal(eng.pred_cls.cpu().numpy().reshape(-1), z["out/pred_classes"].reshape(-1))
    # layers stored in full
    checked = []
    for key in z.files:
        if not key.startswith("layer/"):
            continue
        n = key[len("layer/"):]
        try:
            v = eng.fetch(n)
        except KeyError:
            continue                                       # fused away by the lowering (not materialised)
        e = _rel(v.reshape(z[key].shape), z[key])
        assert e < F32_TOL, (n, e)
        checked.append(n)
    assert meta["seg"] in checked and len(checked) >= 3, checked
    # gradients: stored in full -> element-wise criterion of the oracle test; all of them -> their norms
    grads = eng.get_grads()
    names = meta["train_op_vars"]
    assert set(grads.keys()) == set(names)
    bad = []
    for i, n in enumerate(names):
        ref_norm = float(z["grad_stats"][i][1])
        if ref_norm <= 1e-12:
            continue
        if "grad/" + n in z.files:
            e, fl = _rel2(grads[n], z["grad/" + n]), _rel2(r32["grads"][n], z["grad/" + n])
            if e > F32_TOL + 10 * fl:
                bad.append((n, e, fl))
        e_norm = abs(float(np.linalg.norm(grads[n].astype(np.float64))) - ref_norm) / ref_norm
        fl_norm = abs(float(np.linalg.norm(r32["grads"][n].astype(np.float64))) - ref_norm) / ref_norm
        if e_norm > 10 * F32_TOL + 10 * fl_norm:
            bad.append((n, "norm", e_norm, fl_norm))
    assert not bad, bad[:5]
    print("%s: CUDA f32 vs reference code: logits %.2e, class logits %.2e, %d layers, %d gradients"
          % (snapshot, e_seg, e_cls, len(checked), len(names)))


@pytest.mark.gpu
def test_cuda_cascade_matches_the_reference_code():
    """F4 in f32 mode (tcgen05 split operands at the script's own filter_number = 32) against what the reference's
    unmodified 8AttentionU Train.__init__ computed on its own reader's batch: four sigmoid decoder outputs, four class
    logits, losses, predictions, gradients -- including the pre-ReLU reads of its hand-unrolled trunk."""
    from basi_b200.BAISNet import BAISNet
    from basi_b200.BAISPSPNet import Placeholder
    from basi_b200.BAISRunnerTrain import poly_learning_rate
    from basi_b200.engine import Engine
    meta, z = load("8AttentionU")
    cfg = meta["config"]
    snap = product_snapshot_table()["8AttentionU"]
    S, B = cfg["input_size"][0], cfg["batch_size"]
    params = reference_params(meta)
    net = BAISNet(Placeholder((None, S, S, 4)), is_training=True, num_classes=cfg["num_classes"],
                  num_segment=cfg["num_segment"], segment_attention=cfg["segment_attention"],
                  last_pool_size=cfg["last_pool_size"], filter_number=cfg["filter_number"],
                  attention_module_num=cfg["attention_module_num"])
    eng = Engine(net, B, "f32", True, dict(kind="cascade", pos_weight=snap["pos_weight"],
                                           class_weight=snap["class_weight"]))
    eng.set_params(params)
    lr = float(poly_learning_rate(snap["lr"], float(z["in/step"]), snap["num_steps"]))
    eng.feed(z["in/data"], z["in/label_segment"].astype(np.int32), z["in/label_classes"].astype(np.int32), lr,
             label_att=z["in/label_attention"].astype(np.float32))
    eng.step_device()
    torch.cuda.synchronize()
    r32 = O.attention_u_train_step(params, z["in/data"], z["in/label_segment"], z["in/label_attention"],
                                   z["in/label_classes"], cfg["last_pool_size"], lr, torch.float32, cfg["num_segment"],
                                   cfg["segment_attention"], cfg["attention_module_num"])
    loss, lseg, lcls = eng.losses()
    assert abs(lseg - float(z["out/loss_segment_all"])) < F32_TOL * max(1, abs(float(z["out/loss_segment_all"])))
    assert abs(lcls - float(z["out/loss_class_all"])) < F32_TOL * max(1, abs(float(z["out/loss_class_all"])))
    assert abs(loss - float(z["out/loss"])) < F32_TOL * max(1, abs(float(z["out/loss"])))
    errs = []
    for i in range(4):
        seg = eng.segments[i].t.float().cpu().numpy()
        cl = eng.classes_logits[i].t.float().cpu().numpy().reshape(B, -1)
        e_s = _rel(seg.reshape(z["out/segment_%d" % i].shape), z["out/segment_%d" % i])
        e_c = _rel(cl, z["out/class_%d" % i])
        assert e_s < F32_TOL + 3 * _rel(r32["segments"][i], z["out/segment_%d" % i]), (i, e_s)
        assert e_c < F32_TOL + 3 * _rel(r32["classes"][i], z["out/class_%d" % i]), (i, e_c)
        errs.append((e_s, e_c))
    grads = eng.get_grads()
    names = meta["train_op_vars"]
    assert set(grads.keys()) == set(names)
    bad = []
    for i, n in enumerate(names):
        ref_norm = float(z["grad_stats"][i][1])
        if ref_norm <= 1e-12:
            continue
        if "grad/" + n in z.files:
            e, fl = _rel2(grads[n], z["grad/" + n]), _rel2(r32["grads"][n], z["grad/" + n])
            if e > F32_TOL + 10 * fl:
                bad.append((n, e, fl))
        # every gradient by its norm: a wiring error moves these by O(1); float32 noise does not -- at filter_number 32
        # the smallest |ReLU input| of this batch is 1.7e-7 (oracle.TRACE_RELU_MARGIN), so single mask elements flip in
        # any float32 run and norms move by up to 2.8e-3 (measured on the B200, conv1_1 beta)
        e_norm = abs(float(np.linalg.norm(grads[n].astype(np.float64))) - ref_norm) / ref_norm
        if e_norm > 1e-2:
            bad.append((n, "norm", e_norm))
    assert not bad, bad[:5]
    print("8AttentionU: CUDA f32 vs reference code: (segment, class) errors %s" % (["%.1e/%.1e" % e for e in errs],))
