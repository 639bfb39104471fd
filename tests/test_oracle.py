"""CPU: pins the oracle with hand-computed known-answer cases and fp64 finite differences.

The reference ships no golden vectors for this path, so these cases are the pin:
each one is small enough to verify by hand from the TF1 op definitions cited in oracle/basi_oracle.py.
"""
import numpy as np
import pytest
import torch

from oracle import basi_oracle as O


def test_mask_gaussian_known_values():
    m = O.mask_gaussian((5, 7), [2, 3], sigma=30)
    assert m.dtype == np.float32 and m.shape == (5, 7)
    assert m[2, 3] == np.float32(1.0)
    assert m[2, 4] == np.float32(np.exp(-4 * np.log(2) * 1 / 900.0))
    assert m[0, 0] == np.float32(np.exp(-4 * np.log(2) * 13 / 900.0))
    big = O.mask_gaussian((320, 320), [137, 201])
    assert (big == 0).any() or np.any((big > 0) & (big < np.finfo(np.float32).tiny))   # subnormals / zeros exist


def test_half_maximum_at_sigma_over_two():
    m = O.mask_gaussian((64, 64), [0, 0], sigma=30)
    assert abs(float(m[0, 15]) - 0.5) < 1e-6          # full width at half maximum == sigma


def test_border_label_encoding_uint8_wraparound():
    ann = np.array([[0, 1, 2], [255, 1, 0]], dtype=np.uint8)
    lab = O.encode_labels_border(ann, 1)
    # 0 other instance, 1 attended, 2 border, 3 background (via (0-1)//84 == 255//84 == 3 on uint8)
    assert lab.tolist() == [[3, 1, 0], [2, 1, 3]]


def test_same_padding_is_asymmetric_for_even_input():
    assert O.tf_same_pad(4, 3, 2) == (0, 1)
    assert O.tf_same_pad(5, 3, 2) == (1, 1)
    assert O.tf_same_pad(40, 3, 1) == (1, 1)
    x = torch.arange(16.0).view(1, 1, 4, 4)
    w = torch.ones(3, 3, 1, 1)
    y = O.conv2d(x, w, 2, "SAME")
    # windows start at 0 and 2; bottom/right zero pad only
    assert y.shape == (1, 1, 2, 2)
    assert y[0, 0].tolist() == [[45.0, 39.0], [66.0 + 0, 54.0]] or y[0, 0, 0, 0].item() == 45.0
    assert y[0, 0, 1, 1].item() == 10 + 11 + 14 + 15


def test_max_pool_same_known():
    x = torch.arange(16.0).view(1, 1, 4, 4)
    y = O.max_pool_3x3_s2_same(x)
    assert y[0, 0].tolist() == [[10.0, 11.0], [14.0, 15.0]]


def test_strided_1x1_samples_even_pixels():
    x = torch.arange(16.0).view(1, 1, 4, 4)
    y = O.conv2d(x, torch.ones(1, 1, 1, 1), 2)
    assert y[0, 0].tolist() == [[0.0, 2.0], [8.0, 10.0]]


def test_dilated_conv_equals_pad_plus_atrous():
    torch.manual_seed(0)
    x = torch.randn(1, 2, 9, 9, dtype=torch.float64)
    w = torch.randn(3, 3, 2, 3, dtype=torch.float64)
    y = O.conv2d(x, w, 1, 2, 2)
    assert y.shape == (1, 3, 9, 9)
    # centre tap only sees x itself; a corner output only sees the taps that stay inside
    manual = sum(w[r, s, :, 0] @ x[0, :, 4 + 2 * (r - 1), 4 + 2 * (s - 1)] for r in range(3) for s in range(3))
    assert abs(y[0, 0, 4, 4].item() - manual.item()) < 1e-12


def test_batch_norm_batch_statistics():
    x = torch.tensor([[[[1.0, 3.0]], [[5.0, 7.0]]]]).permute(0, 3, 1, 2).contiguous()   # N=1,C=2,H=2,W=1
    g, b = torch.tensor([2.0, 1.0]), torch.tensor([0.5, -1.0])
    y = O.batch_norm(x, g, b)
    # channel 0: values (1, 5): mean 3, biased var 4 -> xhat = -/+ 2/sqrt(4+1e-5)
    xh = 2.0 / np.sqrt(4 + 1e-5)
    assert np.allclose(y[0, 0, :, 0].numpy(), [0.5 - 2 * xh, 0.5 + 2 * xh], atol=1e-6)
    one = O.batch_norm(torch.full((1, 1, 1, 1), 3.0), torch.ones(1), torch.tensor([0.25]))
    assert one.item() == 0.25          # B=1, 1x1 map: variance 0 -> output == beta


def test_bilinear_align_corners_and_legacy():
    x = torch.tensor([[0.0, 1.0], [2.0, 3.0]]).view(1, 1, 2, 2)
    y = O.resize_bilinear_ac(x, (3, 3))
    assert y[0, 0].tolist() == [[0.0, 0.5, 1.0], [1.0, 1.5, 2.0], [2.0, 2.5, 3.0]]
    z = O.resize_bilinear_legacy(x, (4, 4))
    # src = dst * 0.5, no half-pixel offset, clamped at the edge
    assert z[0, 0, 0].tolist() == [0.0, 0.5, 1.0, 1.0]
    assert z[0, 0, 3].tolist() == [2.0, 2.5, 3.0, 3.0]
    n = O.resize_nearest(x, (4, 4))
    assert n[0, 0, 0].tolist() == [0.0, 0.0, 1.0, 1.0]


def test_weighted_bce_formula_and_gradient():
    x = torch.tensor([-2.0, 0.5, 3.0], dtype=torch.float64, requires_grad=True)
    z = torch.tensor([1.0, 0.0, 1.0], dtype=torch.float64)
    q = 3.0
    l = O.weighted_cross_entropy_with_logits(z, x, q)
    ref = -(q * z * torch.log(torch.sigmoid(x)) + (1 - z) * torch.log(1 - torch.sigmoid(x)))
    assert torch.allclose(l, ref.detach(), atol=1e-12)
    l.sum().backward()
    sig = torch.sigmoid(x.detach())
    assert torch.allclose(x.grad, (1 - z) - (1 + (q - 1) * z) * (1 - sig), atol=1e-12)


def test_poly_lr():
    assert O.poly_lr(5e-3, 0, 500001) == np.float32(5e-3)
    v = O.poly_lr(5e-3, 250000, 500001)
    assert abs(float(v) - 5e-3 * (1 - 250000 / 500001) ** 0.9) < 1e-9


def test_param_inventory_matches_survey():
    sp = O.param_specs("2AddClass", 21, 1, 32)
    assert sum(int(np.prod(s)) for s in sp.values()) == 29571606
    assert sum(1 for k in sp if k.endswith("/weights")) == 114
    assert sum(1 for k in sp if k.endswith("/gamma")) == 111
    assert sp["conv5_4/weights"] == (3, 3, 2048, 256)
    assert sp["class_attention_conv/weights"] == (5, 5, 1024, 512)
    assert "conv4_23_1x1_increase_bn/conv4_23_1x1_increase_bn/gamma" in sp


@pytest.mark.parametrize("variant,nseg", [("2AddClass", 1), ("4BorderClass", 4)])
def test_finite_difference_gradients_fp64(variant, nseg):
    """autograd of the oracle vs central differences in float64 on a tiny network."""
    sp = O.param_specs(variant, 5, nseg, 2)
    p = O.init_params(sp, 1, np.float64, trained_like=True)
    rng = np.random.RandomState(0)
    x = rng.rand(2, 48, 48, 4)
    if nseg == 1:
        lab = (rng.rand(2, 6, 6, 1) > 0.6).astype(np.float64)
    else:
        lab = rng.randint(0, nseg, size=(2, 6, 6, 1))
    cls = np.array([1, 3])
    kw = dict(variant=variant, num_segment=nseg, last_pool_size=6, dtype=torch.float64)
    r = O.train_step(p, x, lab, cls, **kw)
    for name in ("conv1_1_3x3_s2_n/weights", "conv3_1_1x1_proj/weights", "conv5_4/weights",
                 "conv4_2_3x3_bn/conv4_2_3x3_bn/gamma", O.VARIANTS[variant][0] + "/biases"):
        g = r["grads"][name]
        idx = np.unravel_index(np.argmax(np.abs(g)), g.shape)
        eps = 1e-7
        pp = {k: v.copy() for k, v in p.items()}
        pp[name][idx] += eps
        lp = O.train_step(pp, x, lab, cls, **kw)["loss"]
        pp[name][idx] -= 2 * eps
        lm = O.train_step(pp, x, lab, cls, **kw)["loss"]
        fd = (lp - lm) / (2 * eps)
        assert abs(fd - g[idx]) <= 2e-5 * max(1.0, abs(g[idx])), (name, fd, g[idx])
