"""Extracts the golden vectors the REFERENCE's own (vendored TF-slim) tests hold for ops on the hot path and writes them
to tests/golden/slim_reference_tests.json.  Run in the build container (reads /root/reference, nothing else does):

    python tests/golden/make_slim_golden.py

What is taken (parsed with `ast` from the test sources -- literals only, no TensorFlow needed):

* slim/nets/resnet_v1_test.py, class ResnetUtilsTest (the same four tests are repeated in resnet_v2_test.py and are
  checked to agree):
    - testSubsampleThreeByThree / testSubsampleFourByFour: resnet_utils.subsample(x, 2) of range(9) / range(16)
      (= max_pool2d 1x1 stride 2 = x[:, ::2, ::2, :]; the strided 1x1 convolutions of the trunk)
    - testConv2DSameEven / testConv2DSameOdd: input x[h, w] = h + w, kernel w[r, s] = r + s (create_test_input);
      y1 = slim.conv2d 3x3 stride 1 'SAME'; y2 = subsample(y1, 2); y3 = conv2d_same stride 2 (explicit padding, then
      'VALID') == y2; y4 = slim.conv2d 3x3 stride 2 'SAME' -- on the EVEN input TF pads (0 before, 1 after), so y4 != y2.
      These pin the TF 'SAME' / explicit-padding semantics of Network.conv (back/2AddClass/BAISPSPNet.py:118-146),
      in particular of conv1_1_3x3_s2.
* slim/nets/vgg_test.py, class VGG16Test: testEndPoints expected_names and testModelVariables expected_names -- the
  naming contract of the vgg_16 trunk variant B builds on (slim/nets/vgg.py:187-196).
"""
import ast
import json
import os

REF = "/root/reference/slim/nets"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "slim_reference_tests.json")


def _methods(path, cls):
    tree = ast.parse(open(path).read())
    for node in tree.body:
        if isinstance(node, ast.ClassDef) and node.name == cls:
            return {f.name: f for f in node.body if isinstance(f, ast.FunctionDef)}
    raise KeyError(cls)


def _assigned(fn, name):
    """all values assigned to `name` inside the function, in source order (AST nodes)"""
    out = []
    for node in ast.walk(fn):
        if isinstance(node, ast.Assign) and len(node.targets) == 1:
            t = node.targets[0]
            if isinstance(t, ast.Name) and t.id == name:
                out.append(node)
            elif isinstance(t, ast.Tuple) and name in [e.id for e in t.elts if isinstance(e, ast.Name)]:
                out.append(node)
    return sorted(out, key=lambda n: n.lineno)


def _first_list_literal(node):
    """the first list literal below `node` (tf.to_float([[...]]), tf.constant([...]))"""
    for sub in ast.walk(node):
        if isinstance(sub, ast.List):
            try:
                return ast.literal_eval(sub)
            except ValueError:
                continue
    raise ValueError("no literal")


def conv_same(fn):
    n, n2 = ast.literal_eval(_assigned(fn, "n")[0].value)
    res = dict(n=n, n2=n2)
    for name in ("y1_expected", "y2_expected", "y3_expected", "y4_expected"):
        first = _assigned(fn, name)[0].value
        if isinstance(first, ast.Name):          # y3_expected = y2_expected
            res[name] = res[first.id]
        else:
            res[name] = _first_list_literal(first)
    return res


def subsample(fn):
    x = _assigned(fn, "x")[0].value              # tf.reshape(tf.to_float(tf.range(9)), [1, 3, 3, 1])
    count = [ast.literal_eval(c.args[0]) for c in ast.walk(x)
             if isinstance(c, ast.Call) and getattr(c.func, "attr", "") == "range"][0]
    shape = ast.literal_eval(x.args[1])
    e = _assigned(fn, "expected")[0].value
    return dict(range=count, shape=shape, factor=2, expected=_first_list_literal(e.args[0]), expected_shape=ast.literal_eval(e.args[1]))


def main():
    out = {"source": "parsed from /root/reference/slim/nets/{resnet_v1_test,resnet_v2_test,vgg_test}.py by "
                     "tests/golden/make_slim_golden.py"}
    per = []
    for f in ("resnet_v1_test.py", "resnet_v2_test.py"):
        m = _methods(os.path.join(REF, f), "ResnetUtilsTest")
        per.append({
            "subsample_3x3": subsample(m["testSubsampleThreeByThree"]),
            "subsample_4x4": subsample(m["testSubsampleFourByFour"]),
            "conv2d_same_even": conv_same(m["testConv2DSameEven"]),
            "conv2d_same_odd": conv_same(m["testConv2DSameOdd"]),
        })
    assert per[0] == per[1], "resnet_v1_test and resnet_v2_test disagree"
    out["resnet_utils"] = per[0]
    out["resnet_utils"]["cites"] = {
        "subsample_3x3": "slim/nets/resnet_v1_test.py:58-63", "subsample_4x4": "slim/nets/resnet_v1_test.py:65-70",
        "conv2d_same_even": "slim/nets/resnet_v1_test.py:72-112", "conv2d_same_odd": "slim/nets/resnet_v1_test.py:114-153",
        "input": "create_test_input: x[h, w] = h + w (slim/nets/resnet_v1_test.py:30-53); kernel = the same mesh, 3x3"}
    m = _methods(os.path.join(REF, "vgg_test.py"), "VGG16Test")
    out["vgg_16"] = {
        "end_points": _first_list_literal(_assigned(m["testEndPoints"], "expected_names")[0].value),
        "model_variables": _first_list_literal(_assigned(m["testModelVariables"], "expected_names")[0].value),
        "cites": {"end_points": "slim/nets/vgg_test.py:230-259", "model_variables": "slim/nets/vgg_test.py:292-333"}}
    with open(OUT, "w") as f:
        json.dump(out, f, indent=1, sort_keys=True)
    print("wrote", OUT)


if __name__ == "__main__":
    main()
