"""GPU parity of each C-ABI kernel against the oracle (same seeded inputs).  Run with -m gpu on a B200."""
import ctypes as C

import numpy as np
import pytest
import torch

from oracle import basi_oracle as O

pytestmark = pytest.mark.gpu

F32_TOL = 2e-5     # relative to the tensor max-norm (fp32 kernels, different summation order)
BF16_TOL = 1.5e-2  # bf16 storage of inputs/outputs, fp32 accumulate


def _u(rng, *shape):
    return rng.uniform(-1, 1, shape).astype(np.float32)


def nchw(x):
    return torch.from_numpy(np.ascontiguousarray(x)).permute(0, 3, 1, 2)


def nhwc(t):
    return t.permute(0, 2, 3, 1).contiguous().numpy()


# ------------------------------------------------------------------ click map (bit exact)
@pytest.mark.parametrize("B,H,W,sigma", [(3, 320, 320, 30), (2, 64, 48, 20), (1, 8, 8, 30)])
def test_clickmap_pack_bit_exact(B, H, W, sigma):
    from gpu_util import call, dev, host
    from basi_b200.BAISData import click_lut
    rng = np.random.RandomState(0)
    img = rng.randint(0, 256, size=(B, H, W, 3), dtype=np.uint8)
    clicks = np.stack([rng.randint(0, H, B), rng.randint(0, W, B)], 1).astype(np.int32)
    clicks[0] = [H - 1, 0]
    lut = click_lut((H, W), sigma)
    out = torch.zeros((B, H, W, 4), dtype=torch.float32, device="cuda:0")
    imgd, clickd, lutd = dev(img), dev(clicks), dev(lut)
    call("basi_clickmap_pack", imgd.data_ptr(), 0, clickd.data_ptr(), lutd.data_ptr(),
         C.c_int64(lut.size), out.data_ptr(), B, H, W)
    got = host(out)
    for b in range(B):
        ref = O.pack_input(img[b], clicks[b], sigma)
        assert np.array_equal(got[b].view(np.uint32), ref.view(np.uint32))
    # float-image entry (reference feeds /255 float images)
    imgfd = dev(img.astype(np.float32) / 255)
    call("basi_clickmap_pack", imgfd.data_ptr(), 1, clickd.data_ptr(), lutd.data_ptr(),
         C.c_int64(lut.size), out.data_ptr(), B, H, W)
    assert np.array_equal(host(out).view(np.uint32), got.view(np.uint32))


# ------------------------------------------------------------------ convolutions
CONVS = [
    # k, stride, dil, padding, Cin, Cout, H, W, B, bias, relu
    (3, 2, 1, "SAME", 4, 32, 32, 32, 2, False, False),     # conv1_1 (asymmetric SAME pad)
    (3, 2, 1, "SAME", 4, 8, 15, 17, 1, False, False),      # odd input -> symmetric SAME pad
    (3, 1, 1, "SAME", 32, 64, 16, 16, 2, False, False),    # conv1_3
    (1, 1, 1, "VALID", 64, 32, 16, 16, 2, False, False),   # reduce
    (3, 1, 1, 1, 32, 32, 16, 16, 2, False, False),         # zero_padding(1) + 3x3 VALID
    (1, 2, 1, "VALID", 128, 256, 16, 16, 2, False, False), # conv3_1 proj / reduce, stride 2
    (3, 1, 2, 2, 128, 128, 16, 16, 2, False, False),       # conv4 dilated
    (3, 1, 4, 4, 64, 64, 12, 12, 1, False, False),         # conv5 dilated
    (1, 1, 1, "VALID", 256, 1, 8, 8, 2, True, False),      # conv6_n head
    (1, 1, 1, "VALID", 64, 4, 8, 8, 2, True, False),       # conv6_n_4 head
    (5, 5, 1, "VALID", 16, 32, 8, 8, 2, True, True),       # class_attention_conv on a non-5x5 map
    (3, 1, 1, "SAME", 8, 12, 9, 7, 3, False, False),       # ragged channels -> scalar paths
    (1, 1, 1, "VALID", 16, 16, 1, 1, 2, False, False),     # PSP pool1 branch, 1x1 map
]


def _conv_case(case, dtype):
    from gpu_util import act, bf16_round, call, dev, empty_act, host
    from basi_b200._lib import ConvDesc
    k, s, d, padding, cin, cout, H, W, B, bias, relu = case
    rng = np.random.RandomState(hash(case) % 1000)
    x = _u(rng, B, H, W, cin)
    w = (_u(rng, k, k, cin, cout) / np.sqrt(k * k * cin)).astype(np.float32)
    b = _u(rng, cout) if bias else None
    tdt = torch.float32 if dtype == "f32" else torch.bfloat16
    if dtype == "bf16":
        x = bf16_round(x)
    xt = nchw(x).double().requires_grad_(True)
    wt = torch.from_numpy(w).double().requires_grad_(True)
    bt = torch.from_numpy(b).double().requires_grad_(True) if bias else None
    y_ref = O.conv2d(xt, wt, s, padding, d, bt)
    if relu:
        y_ref = torch.relu(y_ref)
    oh, ow = y_ref.shape[2], y_ref.shape[3]
    if padding == "SAME":
        pt, pl = O.tf_same_pad(H, k, s, d)[0], O.tf_same_pad(W, k, s, d)[0]
    elif padding == "VALID":
        pt = pl = 0
    else:
        pt = pl = padding
    desc = ConvDesc(k, k, s, d, pt, pl, 1 if relu else 0)
    out_dt = torch.float32 if bias else tdt
    xa, ya = act(x, tdt), empty_act((B, oh, ow, cout), out_dt, fill=7.0)
    wd = dev(w)
    bd = dev(b) if bias else None
    call("basi_conv_fprop", C.byref(desc), xa.ref, wd.data_ptr(), bd.data_ptr() if bias else None, ya.ref)
    y = host(ya)
    # backward
    dy = _u(rng, B, oh, ow, cout)
    if out_dt == torch.bfloat16:
        dy = bf16_round(dy)
    (y_ref * nchw(dy).double()).sum().backward()
    dya = act(dy, out_dt)
    if relu:
        call("basi_relu_bwd_f32", dya.t.data_ptr(), ya.t.data_ptr(), C.c_int64(dya.t.numel()))
    dxa = empty_act((B, H, W, cin), tdt, fill=3.0)
    call("basi_conv_dgrad", C.byref(desc), dya.ref, wd.data_ptr(), dxa.ref, 0)
    dx0 = host(dxa)
    call("basi_conv_dgrad", C.byref(desc), dya.ref, wd.data_ptr(), dxa.ref, 1)
    dx1 = host(dxa)
    dw = torch.zeros(k, k, cin, cout, device="cuda:0")
    db = torch.zeros(cout, device="cuda:0") if bias else None
    call("basi_conv_wgrad", C.byref(desc), xa.ref, dya.ref, dw.data_ptr(), db.data_ptr() if bias else None)
    return dict(y=y, y_ref=nhwc(y_ref.detach()), dx=dx0, dx2=dx1, dx_ref=nhwc(xt.grad), dw=host(dw),
                dw_ref=wt.grad.numpy(), db=host(db) if bias else None, db_ref=bt.grad.numpy() if bias else None)


@pytest.mark.parametrize("case", CONVS, ids=lambda c: "k%ds%dd%d_%s_%dto%d_%dx%d" % (c[0], c[1], c[2], c[3], c[4], c[5], c[6], c[7]))
@pytest.mark.parametrize("dtype", ["f32", "bf16"])
def test_conv_fprop_dgrad_wgrad(case, dtype):
    from gpu_util import rel_err
    r = _conv_case(case, dtype)
    tol = F32_TOL if dtype == "f32" else BF16_TOL
    assert rel_err(r["y"], r["y_ref"]) < tol
    assert rel_err(r["dx"], r["dx_ref"]) < tol
    assert rel_err(r["dx2"], 2 * r["dx_ref"]) < 2 * tol        # accumulate=1 adds onto the first result
    assert rel_err(r["dw"], r["dw_ref"]) < (F32_TOL * 5 if dtype == "f32" else tol)
    if r["db"] is not None:
        assert rel_err(r["db"], r["db_ref"]) < 1e-4


# ------------------------------------------------------------------ batch norm (+relu, +residual) forward/backward
@pytest.mark.parametrize("dtype", ["f32", "bf16"])
@pytest.mark.parametrize("mode", ["plain", "relu", "res", "res_bn", "res_norelu"])
@pytest.mark.parametrize("shape", [(2, 16, 16, 32), (1, 1, 1, 16), (3, 5, 7, 8), (2, 8, 8, 2048), (4, 40, 40, 128), (3, 37, 41, 512)])
def test_batch_norm_forward_backward(dtype, mode, shape):
    # res_norelu: BN(x) + residual WITHOUT the ReLU -- the materialised junction sum of the 8AttentionU trunk wiring
    from gpu_util import act, bf16_round, call, dev, empty_act, host, rel_err, rel_l2
    B, H, W, Cc = shape
    rng = np.random.RandomState(1)
    tdt = torch.float32 if dtype == "f32" else torch.bfloat16
    rnd = (lambda a: a) if dtype == "f32" else bf16_round
    x = rnd(_u(rng, *shape) * 2 + 0.5)
    x2 = rnd(_u(rng, *shape))
    gamma, beta = rng.uniform(0.5, 1.5, Cc).astype(np.float32), _u(rng, Cc)
    gamma2, beta2 = rng.uniform(0.5, 1.5, Cc).astype(np.float32), _u(rng, Cc)
    dout = rnd(_u(rng, *shape))
    relu = mode not in ("plain", "res_norelu")
    # oracle (float64)
    xt, x2t = nchw(x).double().requires_grad_(True), nchw(x2).double().requires_grad_(True)
    g1, b1 = torch.from_numpy(gamma).double().requires_grad_(True), torch.from_numpy(beta).double().requires_grad_(True)
    g2, b2 = torch.from_numpy(gamma2).double().requires_grad_(True), torch.from_numpy(beta2).double().requires_grad_(True)
    y = O.batch_norm(xt, g1, b1)
    if mode in ("res", "res_norelu"):
        y = y + x2t
    if mode == "res_bn":
        y = y + O.batch_norm(x2t, g2, b2)
    if relu:
        y = torch.relu(y)
    # device forward
    R = float(B * H * W)
    xa, x2a, outa = act(x, tdt), act(x2, tdt), empty_act(shape, tdt)
    sums = torch.zeros(4 * Cc * 8, dtype=torch.float64, device="cuda:0")
    bnp, bnp2 = torch.zeros(4 * Cc, device="cuda:0"), torch.zeros(4 * Cc, device="cuda:0")
    gd, bd, g2d, b2d = dev(gamma), dev(beta), dev(gamma2), dev(beta2)
    cnt = torch.zeros(8, dtype=torch.int32, device="cuda:0")
    call("basi_bn_stats", xa.ref, sums.data_ptr(), gd.data_ptr(), bd.data_ptr(), C.c_double(R), C.c_float(1e-5),
         bnp.data_ptr(), cnt.data_ptr())
    res_ref, res_bnp = None, None
    if mode in ("res", "res_bn", "res_norelu"):
        res_ref = x2a.ref
    if mode == "res_bn":
        call("basi_bn_stats", x2a.ref, sums.data_ptr() + 8 * 2 * Cc * 8, None, None, C.c_double(R), C.c_float(1e-5), None,
             cnt.data_ptr() + 4)                           # unfused form: separate finalize launch
        call("basi_bn_finalize", sums.data_ptr() + 8 * 2 * Cc * 8, g2d.data_ptr(), b2d.data_ptr(), C.c_double(R),
             C.c_float(1e-5), bnp2.data_ptr(), Cc)
        res_bnp = bnp2.data_ptr()
    call("basi_bn_apply", xa.ref, bnp.data_ptr(), res_ref, res_bnp, 1 if relu else 0, outa.ref)
    tol = 1e-5 if dtype == "f32" else 1e-2
    if B * H * W > 1:
        assert rel_err(host(outa), nhwc(y.detach())) < tol
    else:
        assert np.allclose(host(outa), nhwc(y.detach()), atol=1e-2 if dtype == "bf16" else 1e-6)
    # backward against autograd, with the device's own (possibly bf16-rounded) output as the ReLU mask
    if B * H * W == 1:
        return
    (y * nchw(dout).double()).sum().backward()
    da = act(dout, tdt)
    dsums = torch.zeros(2 * Cc * 8, dtype=torch.float64, device="cuda:0")
    coef = torch.zeros(2 * Cc, device="cuda:0")
    dgamma, dbeta = torch.zeros(Cc, device="cuda:0"), torch.zeros(Cc, device="cuda:0")
    dxa = empty_act(shape, tdt, fill=5.0)
    dresa = empty_act(shape, tdt, fill=1.0)
    # plain BN+ReLU recomputes the mask from x; the junction forms read the stored output
    mask = outa.ref if (relu and mode != "relu") else None
    from_x = 1 if mode == "relu" else 0
    call("basi_bn_bwd_reduce", da.ref, mask, xa.ref, bnp.data_ptr(), from_x, dsums.data_ptr(), C.c_double(R),
         dgamma.data_ptr(), dbeta.data_ptr(), coef.data_ptr(), cnt.data_ptr() + 8)
    call("basi_bn_bwd_apply", da.ref, mask, xa.ref, bnp.data_ptr(), coef.data_ptr(), from_x, dxa.ref,
         dresa.ref if mode in ("res", "res_norelu") else None, 1)
    btol = 2e-4 if dtype == "f32" else 3e-2
    assert rel_l2(host(dxa), nhwc(xt.grad)) < btol
    assert rel_err(host(dgamma), g1.grad.numpy()) < btol
    assert rel_err(host(dbeta), b1.grad.numpy()) < btol
    if mode in ("res", "res_norelu"):
        assert rel_l2(host(dresa) - 1.0, nhwc(x2t.grad)) < btol
    if mode == "res_bn":
        dsums.zero_()
        dx2a = empty_act(shape, tdt)
        dg2, db2 = torch.zeros(Cc, device="cuda:0"), torch.zeros(Cc, device="cuda:0")
        call("basi_bn_bwd_reduce", da.ref, mask, x2a.ref, bnp2.data_ptr(), 0, dsums.data_ptr(), C.c_double(R), None,
             None, None, cnt.data_ptr() + 12)              # unfused form
        call("basi_bn_bwd_finalize", dsums.data_ptr(), C.c_double(R), dg2.data_ptr(), db2.data_ptr(),
             coef.data_ptr(), Cc)
        call("basi_bn_bwd_apply", da.ref, mask, x2a.ref, bnp2.data_ptr(), coef.data_ptr(), 0, dx2a.ref, None, 0)
        assert rel_l2(host(dx2a), nhwc(x2t.grad)) < btol
        assert rel_err(host(dg2), g2.grad.numpy()) < btol


@pytest.mark.parametrize("dtype", ["f32", "bf16"])
@pytest.mark.parametrize("shape", [(2, 10, 10, 32), (3, 5, 7, 8), (4, 40, 40, 128)])
def test_relu_of_materialised_sum(dtype, shape):
    """basi_relu_fwd / basi_relu_bwd (8AttentionU trunk wiring: the junction sum and its ReLU are both read): bit-exact
    against numpy, with and without accumulation, also on a strided (channel-slice) destination."""
    from gpu_util import act, bf16_round, call, empty_act, host
    from basi_b200.engine import Act
    rng = np.random.RandomState(4)
    tdt = torch.float32 if dtype == "f32" else torch.bfloat16
    rnd = (lambda a: a.astype(np.float32)) if dtype == "f32" else bf16_round
    x = rnd(_u(rng, *shape))
    x[0, 0, 0, :4] = [0.0, -0.0, 1.0, -1.0]
    dy, old = rnd(_u(rng, *shape)), rnd(_u(rng, *shape))
    xa, ya = act(x, tdt), empty_act(shape, tdt, fill=7.0)
    call("basi_relu_fwd", xa.ref, ya.ref)
    y = np.maximum(x, 0.0)
    assert np.array_equal(host(ya), y)
    for acc in (0, 1):
        dxa = act(old, tdt)
        call("basi_relu_bwd", act(dy, tdt).ref, ya.ref, dxa.ref, acc)
        want = np.where(y > 0, dy, 0.0) + (old if acc else 0.0)
        assert np.array_equal(host(dxa), rnd(want.astype(np.float32)))
    # destination = channel slice of a wider buffer (ld > c)
    B, H, W, Cc = shape
    wide = torch.full((B, H, W, 2 * Cc), 3.0, dtype=tdt, device="cuda:0")
    call("basi_relu_fwd", xa.ref, Act(wide[..., Cc:]).ref)
    torch.cuda.synchronize()
    got = wide.float().cpu().numpy()
    assert np.array_equal(got[..., Cc:], y) and np.all(got[..., :Cc] == 3.0)


# ------------------------------------------------------------------ pooling / bilinear / gate
@pytest.mark.parametrize("dtype", ["f32", "bf16"])
@pytest.mark.parametrize("shape", [(2, 16, 16, 32), (1, 15, 17, 8), (2, 160, 160, 64)])
def test_maxpool_3x3_s2_same(dtype, shape):
    from gpu_util import act, bf16_round, call, empty_act, host, rel_err
    B, H, W, Cc = shape
    rng = np.random.RandomState(2)
    tdt = torch.float32 if dtype == "f32" else torch.bfloat16
    x = _u(rng, *shape)
    if dtype == "bf16":
        x = bf16_round(x)
    xt = nchw(x).double().requires_grad_(True)
    y = O.max_pool_3x3_s2_same(xt)
    oh, ow = y.shape[2], y.shape[3]
    xa, ya = act(x, tdt), empty_act((B, oh, ow, Cc), tdt)
    amax = torch.zeros(B * oh * ow * Cc, dtype=torch.uint8, device="cuda:0")
    call("basi_maxpool3s2_fwd", xa.ref, ya.ref, amax.data_ptr())
    assert np.array_equal(host(ya), nhwc(y.detach()).astype(np.float32))
    dy = _u(rng, B, oh, ow, Cc)
    if dtype == "bf16":
        dy = bf16_round(dy)
    (y * nchw(dy).double()).sum().backward()
    dxa = empty_act(shape, tdt, fill=9.0)
    dya = act(dy, tdt)
    call("basi_maxpool3s2_bwd", dya.ref, amax.data_ptr(), dxa.ref, 0)
    assert rel_err(host(dxa), nhwc(xt.grad)) < (1e-6 if dtype == "f32" else 1e-2)


@pytest.mark.parametrize("dtype", ["f32", "bf16"])
@pytest.mark.parametrize("H,k,Cc", [(40, 40, 64), (40, 20, 64), (40, 13, 64), (40, 6, 64), (40, 8, 1024), (8, 1, 16)])
def test_avgpool(dtype, H, k, Cc):
    from gpu_util import act, bf16_round, call, empty_act, host, rel_err
    B = 2
    rng = np.random.RandomState(3)
    tdt = torch.float32 if dtype == "f32" else torch.bfloat16
    x = _u(rng, B, H, H, Cc)
    if dtype == "bf16":
        x = bf16_round(x)
    xt = nchw(x).double().requires_grad_(True)
    y = O.avg_pool(xt, k)
    o = y.shape[2]
    xa, ya = act(x, tdt), empty_act((B, o, o, Cc), tdt)
    call("basi_avgpool_fwd", xa.ref, k, ya.ref)
    assert rel_err(host(ya), nhwc(y.detach())) < (1e-5 if dtype == "f32" else 1e-2)
    dy = _u(rng, B, o, o, Cc)
    if dtype == "bf16":
        dy = bf16_round(dy)
    (y * nchw(dy).double()).sum().backward()
    dxa = empty_act((B, H, H, Cc), tdt, fill=2.0)
    dya = act(dy, tdt)
    call("basi_avgpool_bwd", dya.ref, k, dxa.ref, 0)
    assert rel_err(host(dxa), nhwc(xt.grad)) < (1e-5 if dtype == "f32" else 1e-2)
    call("basi_avgpool_bwd", dya.ref, k, dxa.ref, 1)
    assert rel_err(host(dxa), 2 * nhwc(xt.grad)) < (1e-5 if dtype == "f32" else 2e-2)


@pytest.mark.parametrize("dtype", ["f32", "bf16"])
@pytest.mark.parametrize("o,P", [(1, 8), (2, 8), (3, 40), (6, 40), (4, 8)])
def test_bilinear_align_corners(dtype, o, P):
    from gpu_util import act, bf16_round, call, empty_act, host, rel_err
    B, Cc = 2, 32
    rng = np.random.RandomState(4)
    tdt = torch.float32 if dtype == "f32" else torch.bfloat16
    x = _u(rng, B, o, o, Cc)
    if dtype == "bf16":
        x = bf16_round(x)
    xt = nchw(x).double().requires_grad_(True)
    y = O.resize_bilinear_ac(xt, (P, P))
    # write into a channel slice of a wider buffer, like the PSP concat
    wide = torch.zeros((B, P, P, 3 * Cc), dtype=tdt, device="cuda:0")
    from basi_b200.engine import Act
    ya = Act(wide[..., Cc:2 * Cc])
    xa = act(x, tdt)
    call("basi_bilinear_ac_fwd", xa.ref, ya.ref)
    got = host(wide)
    assert rel_err(got[..., Cc:2 * Cc], nhwc(y.detach())) < (1e-5 if dtype == "f32" else 1e-2)
    assert np.all(got[..., :Cc] == 0) and np.all(got[..., 2 * Cc:] == 0)
    dy = _u(rng, B, P, P, 3 * Cc)
    if dtype == "bf16":
        dy = bf16_round(dy)
    (y * nchw(dy[..., Cc:2 * Cc]).double()).sum().backward()
    dwide = act(dy, tdt)
    dxa = empty_act((B, o, o, Cc), tdt, fill=4.0)
    dslice = Act(dwide.t[..., Cc:2 * Cc])
    call("basi_bilinear_ac_bwd", dslice.ref, dxa.ref, 0)
    assert rel_err(host(dxa), nhwc(xt.grad)) < (2e-5 if dtype == "f32" else 1e-2)


@pytest.mark.parametrize("dtype", ["f32", "bf16"])
@pytest.mark.parametrize("nseg,att", [(1, 0), (4, 1), (3, 2)])
def test_gate_multiply(dtype, nseg, att):
    from gpu_util import act, bf16_round, call, dev, empty_act, host, rel_err
    B, P, Cc = 2, 8, 64
    rng = np.random.RandomState(5)
    tdt = torch.float32 if dtype == "f32" else torch.bfloat16
    rnd = (lambda a: a) if dtype == "f32" else bf16_round
    feat, logits, dout = rnd(_u(rng, B, P, P, Cc)), _u(rng, B, P, P, nseg) * 3, rnd(_u(rng, B, P, P, Cc))
    ft = torch.from_numpy(feat).double().requires_grad_(True)
    lt = torch.from_numpy(logits).double().requires_grad_(True)
    y = torch.relu(ft) * lt[..., att:att + 1]
    (y * torch.from_numpy(dout).double()).sum().backward()
    fa, ya, ld = act(feat, tdt), empty_act((B, P, P, Cc), tdt), dev(logits)
    call("basi_gate_mul_fwd", fa.ref, ld.data_ptr(), nseg, att, ya.ref)
    assert rel_err(host(ya), y.detach().numpy()) < (1e-6 if dtype == "f32" else 1e-2)
    dfa = empty_act((B, P, P, Cc), tdt, fill=1.0)
    dl = torch.full((B, P, P, nseg), 0.5, device="cuda:0")
    douta = act(dout, tdt)
    call("basi_gate_mul_bwd", douta.ref, fa.ref, ld.data_ptr(), nseg, att, dfa.ref, 1, dl.data_ptr())
    assert rel_err(host(dfa) - 1.0, ft.grad.numpy()) < (1e-5 if dtype == "f32" else 2e-2)
    assert rel_err(host(dl) - 0.5, lt.grad.numpy()) < 1e-4


# ------------------------------------------------------------------ losses / SGD / predictions / skinny GEMMs
@pytest.mark.parametrize("n", [1, 777, 25600])
def test_weighted_bce_fused(n):
    from gpu_util import call, dev, host, rel_err
    rng = np.random.RandomState(6)
    x = (_u(rng, n) * 12).astype(np.float32)
    z = (rng.rand(n) > 0.7).astype(np.float32)
    xt = torch.from_numpy(x).double().requires_grad_(True)
    loss = O.weighted_cross_entropy_with_logits(torch.from_numpy(z).double(), xt, 3.0).mean()
    loss.backward()
    acc = torch.zeros(2, dtype=torch.float64, device="cuda:0")
    dl = torch.zeros(n, device="cuda:0")
    xd, zd = dev(x), dev(z)
    call("basi_wbce_fwd_bwd", xd.data_ptr(), zd.data_ptr(), C.c_float(3.0), C.c_double(1.0 / n),
         C.c_float(1.0 / n), C.c_int64(n), acc.data_ptr(), dl.data_ptr())
    assert abs(host(acc)[0] - loss.item()) < 1e-6 * max(1.0, abs(loss.item()))
    assert rel_err(host(dl), xt.grad.numpy()) < 1e-5


@pytest.mark.parametrize("rows,Cc", [(16, 21), (2, 91), (3200, 4), (128, 3)])
def test_softmax_ce_fused(rows, Cc):
    from gpu_util import call, dev, host, rel_err
    rng = np.random.RandomState(7)
    x = (_u(rng, rows, Cc) * 6).astype(np.float32)
    lab = rng.randint(0, Cc, rows).astype(np.int32)
    xt = torch.from_numpy(x).double().requires_grad_(True)
    loss = torch.nn.functional.cross_entropy(xt, torch.from_numpy(lab).long())
    (0.2 * loss).backward()
    acc = torch.zeros(2, dtype=torch.float64, device="cuda:0")
    dl = torch.zeros(rows, Cc, device="cuda:0")
    xd, labd = dev(x), dev(lab)
    call("basi_softmax_ce_fwd_bwd", xd.data_ptr(), labd.data_ptr(), C.c_int64(rows), Cc,
         C.c_double(1.0 / rows), C.c_float(0.2 / rows), acc.data_ptr(), dl.data_ptr())
    assert abs(host(acc)[0] - loss.item()) < 1e-5 * max(1.0, abs(loss.item()))
    assert rel_err(host(dl), xt.grad.numpy()) < 1e-5


def test_sgd_bit_exact_and_bf16_copy():
    from gpu_util import call, dev, host
    rng = np.random.RandomState(8)
    n = 100003
    w, g = _u(rng, n), _u(rng, n)
    wd, gd, lr = dev(w), dev(g), dev(np.array([5e-3], dtype=np.float32))
    wb = torch.zeros(n, dtype=torch.bfloat16, device="cuda:0")
    call("basi_sgd_step", wd.data_ptr(), gd.data_ptr(), lr.data_ptr(), C.c_int64(n), wb.data_ptr())
    ref = w - np.float32(5e-3) * g
    got = host(wd)
    # fused multiply-add vs separate rounding: allow 1 ulp
    assert np.max(np.abs(got - ref)) <= np.max(np.abs(np.spacing(ref)))
    assert np.array_equal(host(wb), torch.from_numpy(got).to(torch.bfloat16).float().numpy())


def test_predictions_bit_exact():
    from gpu_util import call, dev, host
    rng = np.random.RandomState(9)
    B, P, Cc, S = 2, 8, 4, 64
    logits = (_u(rng, B, P, P, Cc) * 4).astype(np.float32)
    logits[0, 0, 0] = [1.0, 1.0, 0.5, 1.0]                      # tie -> first index wins
    ld = dev(logits)
    out = torch.zeros(B * P * P, dtype=torch.int32, device="cuda:0")
    call("basi_argmax", ld.data_ptr(), C.c_int64(B * P * P), Cc, out.data_ptr())
    assert np.array_equal(host(out).reshape(B, P, P), np.argmax(logits, -1))
    one = (_u(rng, 1000) * 2).astype(np.float32)
    one[:3] = [0.5, 0.50000006, 0.49999997]
    th = torch.zeros(1000, dtype=torch.int32, device="cuda:0")
    oned = dev(one)
    call("basi_threshold", oned.data_ptr(), C.c_float(0.5), th.data_ptr(), C.c_int64(1000))
    assert np.array_equal(host(th), (one > 0.5).astype(np.int32))
    up = torch.zeros(B * S * S, dtype=torch.int32, device="cuda:0")
    call("basi_upsample_legacy_argmax", ld.data_ptr(), B, P, P, Cc, S, S, up.data_ptr())
    ref = O.predict_click(logits, (S, S))
    agree = np.mean(host(up).reshape(B, S, S) == ref)
    assert agree >= 0.999, agree                                 # float32 lerp order may flip exact ties only


@pytest.mark.parametrize("dtype", ["f32", "bf16"])
@pytest.mark.parametrize("M,K,N,relu", [(16, 25600, 512, True), (2, 512, 21, False), (1, 1600, 64, True), (4, 512, 91, False)])
def test_skinny_gemms(dtype, M, K, N, relu):
    from gpu_util import bf16_round, call, dev, host, rel_err
    rng = np.random.RandomState(10)
    tdt = torch.float32 if dtype == "f32" else torch.bfloat16
    code = 0 if dtype == "f32" else 1
    a = _u(rng, M, K)
    if dtype == "bf16":
        a = bf16_round(a)
    w, b, dy = (_u(rng, K, N) / np.sqrt(K)).astype(np.float32), _u(rng, N), _u(rng, M, N)
    at = torch.from_numpy(a).double().requires_grad_(True)
    wt = torch.from_numpy(w).double().requires_grad_(True)
    bt = torch.from_numpy(b).double().requires_grad_(True)
    y = at @ wt + bt
    if relu:
        y = torch.relu(y)
    (y * torch.from_numpy(dy).double()).sum().backward()
    ad, wd, bd = dev(a, tdt), dev(w), dev(b)
    yd = torch.full((M, N), 3.0, device="cuda:0")
    call("basi_skinny_fwd", ad.data_ptr(), code, C.c_int64(K), wd.data_ptr(), bd.data_ptr(), yd.data_ptr(), M, K, N,
         1 if relu else 0)
    assert rel_err(host(yd), y.detach().numpy()) < 2e-5
    # workspace variant: same result, and bit-identical from call to call (fixed summation order)
    from basi_b200 import _lib
    nws = int(_lib.load().basi_skinny_fwd_workspace_floats(M, K, N))
    assert nws >= M * N
    ws = torch.full((nws,), 7.0, device="cuda:0")
    runs = []
    for _ in range(3):
        yw = torch.full((M, N), -3.0, device="cuda:0")
        call("basi_skinny_fwd_ws", ad.data_ptr(), code, C.c_int64(K), wd.data_ptr(), bd.data_ptr(), yw.data_ptr(), M, K,
             N, 1 if relu else 0, ws.data_ptr())
        runs.append(host(yw))
    assert rel_err(runs[0], y.detach().numpy()) < 2e-5
    assert np.array_equal(runs[0], runs[1]) and np.array_equal(runs[0], runs[2])
    dyd = dev(dy)
    if relu:
        call("basi_relu_bwd_f32", dyd.data_ptr(), yd.data_ptr(), C.c_int64(M * N))
    dw, db = torch.zeros(K, N, device="cuda:0"), torch.zeros(N, device="cuda:0")
    call("basi_skinny_wgrad", ad.data_ptr(), code, C.c_int64(K), dyd.data_ptr(), dw.data_ptr(), db.data_ptr(), M, K, N)
    assert rel_err(host(dw), wt.grad.numpy()) < 2e-5
    assert rel_err(host(db), bt.grad.numpy()) < 2e-5
    da = torch.zeros(M, K, dtype=tdt, device="cuda:0")
    call("basi_skinny_dgrad", dyd.data_ptr(), wd.data_ptr(), da.data_ptr(), code, C.c_int64(K), M, K, N, 0)
    assert rel_err(host(da), at.grad.numpy()) < (2e-5 if dtype == "f32" else 1e-2)


@pytest.mark.parametrize("dtype", ["f32", "bf16"])
def test_batch_norm_junction_backward_on_concat_slices(dtype):
    """out / dout of the conv5_3 junction are channel slices of the PSP concat buffer (ld != c)."""
    from gpu_util import act, bf16_round, call, dev, empty_act, host, rel_err
    from basi_b200.engine import Act
    B, H, W, Cc = 1, 40, 40, 256
    shape = (B, H, W, Cc)
    rng = np.random.RandomState(11)
    tdt = torch.float32 if dtype == "f32" else torch.bfloat16
    rnd = (lambda a: a) if dtype == "f32" else bf16_round
    x, res, dout = rnd(_u(rng, *shape) * 2), rnd(_u(rng, *shape)), rnd(_u(rng, *shape))
    gamma, beta = rng.uniform(0.05, 0.25, Cc).astype(np.float32), _u(rng, Cc)
    xt, rt = nchw(x).double().requires_grad_(True), nchw(res).double().requires_grad_(True)
    g1, b1 = torch.from_numpy(gamma).double().requires_grad_(True), torch.from_numpy(beta).double().requires_grad_(True)
    y = torch.relu(O.batch_norm(xt, g1, b1) + rt)
    (y * nchw(dout).double()).sum().backward()
    R = float(B * H * W)
    xa, ra = act(x, tdt), act(res, tdt)
    wide = torch.zeros((B, H, W, 2 * Cc), dtype=tdt, device="cuda:0")
    dwide = torch.zeros((B, H, W, 2 * Cc), dtype=tdt, device="cuda:0")
    dwide[..., :Cc] = torch.from_numpy(dout).to("cuda:0").to(tdt)
    outa, da = Act(wide[..., :Cc]), Act(dwide[..., :Cc])
    sums = torch.zeros(4 * Cc * 8, dtype=torch.float64, device="cuda:0")
    cnt = torch.zeros(4, dtype=torch.int32, device="cuda:0")
    bnp, coef = torch.zeros(4 * Cc, device="cuda:0"), torch.zeros(2 * Cc, device="cuda:0")
    gd, bd = dev(gamma), dev(beta)
    call("basi_bn_stats", xa.ref, sums.data_ptr(), gd.data_ptr(), bd.data_ptr(), C.c_double(R), C.c_float(1e-5),
         bnp.data_ptr(), cnt.data_ptr())
    call("basi_bn_apply", xa.ref, bnp.data_ptr(), ra.ref, None, 1, outa.ref)
    tol = 1e-5 if dtype == "f32" else 1e-2
    assert rel_err(host(wide)[..., :Cc], nhwc(y.detach())) < tol
    dgamma, dbeta = torch.zeros(Cc, device="cuda:0"), torch.zeros(Cc, device="cuda:0")
    dxa, dra = empty_act(shape, tdt, fill=5.0), empty_act(shape, tdt, fill=7.0)
    call("basi_bn_bwd_reduce", da.ref, outa.ref, xa.ref, bnp.data_ptr(), 0, sums.data_ptr() + 16 * Cc * 8, C.c_double(R),
         dgamma.data_ptr(), dbeta.data_ptr(), coef.data_ptr(), cnt.data_ptr() + 4)
    call("basi_bn_bwd_apply", da.ref, outa.ref, xa.ref, bnp.data_ptr(), coef.data_ptr(), 0, dxa.ref, dra.ref, 0)
    btol = 2e-4 if dtype == "f32" else 3e-2
    assert rel_err(host(dbeta), b1.grad.numpy()) < btol
    assert rel_err(host(dgamma), g1.grad.numpy()) < btol
    assert rel_err(host(dra), nhwc(rt.grad)) < btol
    assert rel_err(host(dxa), nhwc(xt.grad)) < btol


# ------------------------------------------------------------------ variant-B gating ops
@pytest.mark.parametrize("dtype", ["f32", "bf16"])
def test_mask_multiply(dtype):
    from gpu_util import act, bf16_round, call, dev, empty_act, host, rel_err
    B, P, Cc = 2, 20, 64
    rng = np.random.RandomState(21)
    tdt = torch.float32 if dtype == "f32" else torch.bfloat16
    rnd = (lambda a: a) if dtype == "f32" else bf16_round
    feat, mask, dout = rnd(_u(rng, B, P, P, Cc)), rng.rand(B, P, P, 1).astype(np.float32), rnd(_u(rng, B, P, P, Cc))
    ft = nchw(feat).double().requires_grad_(True)
    mt = nchw(mask).double().requires_grad_(True)
    y = O.mask_multiply(ft, mt)
    (y * nchw(dout).double()).sum().backward()
    fa, ya, md = act(feat, tdt), empty_act((B, P, P, Cc), tdt), dev(mask)
    call("basi_mask_mul_fwd", fa.ref, md.data_ptr(), 1, 0, ya.ref)
    assert rel_err(host(ya), nhwc(y.detach())) < (1e-6 if dtype == "f32" else 1e-2)
    dfa = empty_act((B, P, P, Cc), tdt, fill=2.0)
    dm = torch.zeros((B, P, P, 1), device="cuda:0")
    douta = act(dout, tdt)
    call("basi_mask_mul_bwd", douta.ref, fa.ref, md.data_ptr(), 1, 0, dfa.ref, 1, dm.data_ptr())
    assert rel_err(host(dfa) - 2.0, nhwc(ft.grad)) < (1e-5 if dtype == "f32" else 2e-2)
    assert rel_err(host(dm), nhwc(mt.grad)) < 1e-4
    call("basi_mask_mul_bwd", douta.ref, fa.ref, md.data_ptr(), 1, 0, dfa.ref, 0, None)     # click map: no gradient
    assert rel_err(host(dfa), nhwc(ft.grad)) < (1e-5 if dtype == "f32" else 2e-2)


@pytest.mark.parametrize("Cc,sel,thr", [(2, 1, 0.9), (2, 1, 0.5), (4, 1, -1.0)])
def test_softmax_attention_gate(Cc, sel, thr):
    from gpu_util import call, dev, host, rel_err
    rng = np.random.RandomState(22)
    B, P = 2, 24
    logits = (_u(rng, B, P, P, Cc) * 5).astype(np.float32)
    dgate = _u(rng, B, P, P, 1)
    lt = nchw(logits).double().requires_grad_(True)
    g = O.attention_gate(lt, sel, thr)
    (g * nchw(dgate).double()).sum().backward()
    ld, dgd = dev(logits), dev(dgate)
    gate = torch.zeros(B * P * P, device="cuda:0")
    rows = B * P * P
    call("basi_softmax_gate_fwd", ld.data_ptr(), C.c_int64(rows), Cc, sel, C.c_float(thr), gate.data_ptr())
    ref = nhwc(g.detach()).reshape(-1)
    got = host(gate)
    assert np.array_equal(got > 0, ref > 0)                       # the hard gate selects the same pixels
    assert rel_err(got, ref) < 1e-5
    dl = torch.full((rows, Cc), 0.25, device="cuda:0")
    call("basi_softmax_gate_bwd", ld.data_ptr(), dgd.data_ptr(), C.c_int64(rows), Cc, sel, C.c_float(thr),
         dl.data_ptr(), 1)
    assert rel_err(host(dl) - 0.25, nhwc(lt.grad).reshape(rows, Cc)) < 1e-4


@pytest.mark.parametrize("dtype", ["f32", "bf16"])
@pytest.mark.parametrize("ih,oh,Cc", [(40, 80, 1), (45, 90, 16), (40, 17, 8), (320, 40, 1), (7, 20, 3)])
def test_resize_nearest_neighbor(dtype, ih, oh, Cc):
    from gpu_util import act, bf16_round, call, empty_act, host, rel_err
    B = 2
    rng = np.random.RandomState(23)
    tdt = torch.float32 if dtype == "f32" else torch.bfloat16
    rnd = (lambda a: a) if dtype == "f32" else bf16_round
    x = rnd(_u(rng, B, ih, ih + 1, Cc))
    ow = oh + 3
    ref = O.resize_nearest(nchw(x), (oh, ow))
    xa, ya = act(x, tdt), empty_act((B, oh, ow, Cc), tdt)
    call("basi_resize_nearest_fwd", xa.ref, ya.ref)
    assert np.array_equal(host(ya), nhwc(ref))                    # pure gather: bit exact
    # adjoint: <resize(x), dy> == <x, resize^T(dy)>
    dy = rnd(_u(rng, B, oh, ow, Cc))
    dya, dxa = act(dy, tdt), empty_act((B, ih, ih + 1, Cc), tdt, fill=0.5)
    call("basi_resize_nearest_bwd", dya.ref, dxa.ref, 1)
    xt = nchw(x).double().requires_grad_(True)
    (O.resize_nearest(xt, (oh, ow)) * nchw(dy).double()).sum().backward()
    assert rel_err(host(dxa) - 0.5, nhwc(xt.grad)) < (1e-6 if dtype == "f32" else 2e-2)


# ------------------------------------------------------------------ strided 1x1 conv = subsample + 1x1 conv (bit exact)
@pytest.mark.parametrize("dtype", ["f32", "bf16"])
@pytest.mark.parametrize("shape,s", [((2, 80, 80, 128), 2), ((1, 7, 9, 8), 2), ((2, 10, 10, 16), 3), ((1, 4, 4, 8), 1)])
def test_subsample_and_adjoint_bit_exact(dtype, shape, s):
    from gpu_util import act, bf16_round, call, empty_act, host
    td = torch.float32 if dtype == "f32" else torch.bfloat16
    rng = np.random.RandomState(5)
    x = _u(rng, *shape)
    if dtype == "bf16":
        x = bf16_round(x)
    B, H, W, Cc = shape
    oh, ow = -(-H // s), -(-W // s)
    xa = act(x, td)
    ya = empty_act((B, oh, ow, Cc), td)
    call("basi_subsample_fwd", xa.ref, s, ya.ref)
    assert np.array_equal(host(ya), x[:, ::s, ::s, :])
    dy = _u(rng, B, oh, ow, Cc)
    if dtype == "bf16":
        dy = bf16_round(dy)
    dya = act(dy, td)
    dxa = empty_act(shape, td, fill=7.0)        # stale contents must be overwritten when accumulate == 0
    call("basi_subsample_bwd", dya.ref, s, dxa.ref, 0)
    want = np.zeros(shape, np.float32)
    want[:, ::s, ::s, :] = dy
    assert np.array_equal(host(dxa), want)
    base = _u(rng, *shape)
    if dtype == "bf16":
        base = bf16_round(base)
    dxb = act(base, td)
    call("basi_subsample_bwd", dya.ref, s, dxb.ref, 1)
    want = base.copy()
    want[:, ::s, ::s, :] += dy
    if dtype == "bf16":
        want = bf16_round(want)
    assert np.array_equal(host(dxb), want)


# ------------------------------------------------------------------ conv1_1 stem kernels (fp32 NHWC4 input, Cout 32 / 64)
@pytest.mark.parametrize("out_dtype", ["f32", "bf16"])
@pytest.mark.parametrize("cout,H,W,s,B", [(32, 64, 64, 2, 2), (64, 33, 47, 2, 1), (32, 17, 9, 1, 3), (32, 320, 320, 2, 2)])
def test_stem_conv_fprop_wgrad(out_dtype, cout, H, W, s, B):
    from gpu_util import act, bf16_round, call, dev, empty_act, host, rel_err
    from basi_b200._lib import ConvDesc
    rng = np.random.RandomState(cout + H)
    x = _u(rng, B, H, W, 4)
    w = (_u(rng, 3, 3, 4, cout) / 6.0).astype(np.float32)
    xt = nchw(x).double()
    wt = torch.from_numpy(w).double().requires_grad_(True)
    y_ref = O.conv2d(xt, wt, s, "SAME", 1, None)
    oh, ow = y_ref.shape[2], y_ref.shape[3]
    pt, pl = O.tf_same_pad(H, 3, s, 1)[0], O.tf_same_pad(W, 3, s, 1)[0]
    desc = ConvDesc(3, 3, s, 1, pt, pl, 0)
    tdt = torch.float32 if out_dtype == "f32" else torch.bfloat16
    xa, ya = act(x), empty_act((B, oh, ow, cout), tdt, fill=7.0)
    wd = dev(w)
    call("basi_conv_fprop", C.byref(desc), xa.ref, wd.data_ptr(), None, ya.ref)
    tol = F32_TOL if out_dtype == "f32" else BF16_TOL
    assert rel_err(host(ya), nhwc(y_ref.detach())) < tol
    dy = _u(rng, B, oh, ow, cout)
    if out_dtype == "bf16":
        dy = bf16_round(dy)
    (y_ref * nchw(dy).double()).sum().backward()
    dya = act(dy, tdt)
    dw = torch.full((3, 3, 4, cout), 0.5, device="cuda:0")          # wgrad accumulates onto dw
    call("basi_conv_wgrad", C.byref(desc), xa.ref, dya.ref, dw.data_ptr(), None)
    assert rel_err(host(dw) - 0.5, wt.grad.numpy()) < 1e-4


# ------------------------------------------------------------------ resident (cooperative) BN backward
@pytest.mark.parametrize("dtype", ["f32", "bf16"])
@pytest.mark.parametrize("relu", [True, False])
@pytest.mark.parametrize("shape", [(16, 40, 40, 128), (16, 40, 40, 256), (4, 80, 80, 32), (3, 37, 41, 64), (2, 33, 35, 8)])
def test_batch_norm_backward_resident(dtype, relu, shape):
    """basi_bn_bwd_fused == basi_bn_bwd_reduce + basi_bn_bwd_apply, and both match autograd of the oracle BN."""
    from gpu_util import act, bf16_round, call, dev, empty_act, host, rel_err, rel_l2
    from basi_b200 import _lib
    B, H, W, Cc = shape
    rng = np.random.RandomState(3)
    tdt = torch.float32 if dtype == "f32" else torch.bfloat16
    rnd = (lambda a: a) if dtype == "f32" else bf16_round
    x = rnd(_u(rng, *shape) * 2 + 0.5)
    dout = rnd(_u(rng, *shape))
    gamma, beta = rng.uniform(0.5, 1.5, Cc).astype(np.float32), _u(rng, Cc)
    xt = nchw(x).double().requires_grad_(True)
    g1, b1 = torch.from_numpy(gamma).double().requires_grad_(True), torch.from_numpy(beta).double().requires_grad_(True)
    y = O.batch_norm(xt, g1, b1)
    if relu:
        y = torch.relu(y)
    (y * nchw(dout).double()).sum().backward()
    R = float(B * H * W)
    xa, da = act(x, tdt), act(dout, tdt)
    if _lib.load().basi_bn_bwd_fused_supported(xa.ref) != 1:
        pytest.skip("tensor does not fit the resident kernel on this device")
    sums = torch.zeros(2 * Cc * 8, dtype=torch.float64, device="cuda:0")
    bnp = torch.zeros(4 * Cc, device="cuda:0")
    gd, bd = dev(gamma), dev(beta)
    cnt = torch.zeros(8, dtype=torch.int32, device="cuda:0")
    call("basi_bn_stats", xa.ref, sums.data_ptr(), gd.data_ptr(), bd.data_ptr(), C.c_double(R), C.c_float(1e-5),
         bnp.data_ptr(), cnt.data_ptr())
    from_x = 1 if relu else 0
    outs = []
    for fused in (False, True, True):                      # twice fused: the barrier word is reused without a reset
        dsums = torch.zeros(2 * Cc * 8, dtype=torch.float64, device="cuda:0")
        coef = torch.zeros(2 * Cc, device="cuda:0")
        dgamma, dbeta = torch.full((Cc,), 0.25, device="cuda:0"), torch.full((Cc,), -0.5, device="cuda:0")
        dxa = empty_act(shape, tdt, fill=5.0)
        if fused:
            call("basi_bn_bwd_fused", da.ref, xa.ref, bnp.data_ptr(), from_x, dsums.data_ptr(), C.c_double(R),
                 dgamma.data_ptr(), dbeta.data_ptr(), coef.data_ptr(), cnt.data_ptr() + 16, dxa.ref)
        else:
            call("basi_bn_bwd_reduce", da.ref, None, xa.ref, bnp.data_ptr(), from_x, dsums.data_ptr(), C.c_double(R),
                 dgamma.data_ptr(), dbeta.data_ptr(), coef.data_ptr(), cnt.data_ptr() + 8)
            call("basi_bn_bwd_apply", da.ref, None, xa.ref, bnp.data_ptr(), coef.data_ptr(), from_x, dxa.ref, None, 0)
        outs.append((host(dxa), host(dgamma) - 0.25, host(dbeta) + 0.5, host(coef)))
    btol = 2e-4 if dtype == "f32" else 3e-2
    for dx, dg, db, _ in outs:
        assert rel_l2(dx, nhwc(xt.grad)) < btol
        assert rel_err(dg, g1.grad.numpy()) < btol
        assert rel_err(db, b1.grad.numpy()) < btol
    # fused vs pair: same arithmetic up to the summation order of the per-channel sums
    ptol = 1e-5 if dtype == "f32" else 1e-2
    for k in (1, 2):
        assert rel_err(outs[k][0], outs[0][0]) < ptol
        assert rel_err(outs[k][1], outs[0][1]) < 1e-5 and rel_err(outs[k][2], outs[0][2]) < 1e-5
        assert rel_err(outs[k][3], outs[0][3]) < 1e-5


# ------------------------------------------------------------------ packed ReLU mask of the residual junctions
@pytest.mark.parametrize("mode", ["res", "res_bn"])
@pytest.mark.parametrize("shape", [(4, 40, 40, 128), (3, 37, 41, 256), (16, 40, 40, 512)])
def test_batch_norm_junction_mask_bits(mode, shape):
    """basi_bn_apply_bits / _bwd_reduce_bits / _bwd_apply_bits == the pair that re-reads the stored output."""
    from gpu_util import act, bf16_round, call, dev, empty_act, host, rel_err
    from basi_b200 import _lib
    B, H, W, Cc = shape
    rng = np.random.RandomState(11)
    tdt = torch.bfloat16
    x, x2 = bf16_round(_u(rng, *shape) * 2 + 0.5), bf16_round(_u(rng, *shape))
    dout, dres0 = bf16_round(_u(rng, *shape)), bf16_round(_u(rng, *shape))
    gamma, beta = rng.uniform(0.5, 1.5, Cc).astype(np.float32), _u(rng, Cc)
    gamma2, beta2 = rng.uniform(0.5, 1.5, Cc).astype(np.float32), _u(rng, Cc)
    R = float(B * H * W)
    xa, x2a, da = act(x, tdt), act(x2, tdt), act(dout, tdt)
    assert _lib.load().basi_bn_maskbits_supported(xa.ref) == 1
    gd, bd, g2d, b2d = dev(gamma), dev(beta), dev(gamma2), dev(beta2)
    sums = torch.zeros(4 * Cc * 8, dtype=torch.float64, device="cuda:0")
    bnp, bnp2 = torch.zeros(4 * Cc, device="cuda:0"), torch.zeros(4 * Cc, device="cuda:0")
    cnt = torch.zeros(16, dtype=torch.int32, device="cuda:0")
    call("basi_bn_stats", xa.ref, sums.data_ptr(), gd.data_ptr(), bd.data_ptr(), C.c_double(R), C.c_float(1e-5),
         bnp.data_ptr(), cnt.data_ptr())
    res_bnp = None
    if mode == "res_bn":
        call("basi_bn_stats", x2a.ref, sums.data_ptr() + 8 * 2 * Cc * 8, g2d.data_ptr(), b2d.data_ptr(), C.c_double(R),
             C.c_float(1e-5), bnp2.data_ptr(), cnt.data_ptr() + 4)
        res_bnp = bnp2.data_ptr()
    out_a, out_b = empty_act(shape, tdt), empty_act(shape, tdt)
    bits = torch.zeros(B * H * W * Cc // 8, dtype=torch.uint8, device="cuda:0")
    call("basi_bn_apply", xa.ref, bnp.data_ptr(), x2a.ref, res_bnp, 1, out_a.ref)
    call("basi_bn_apply_bits", xa.ref, bnp.data_ptr(), x2a.ref, res_bnp, 1, out_b.ref, bits.data_ptr())
    oa = host(out_a)
    assert np.array_equal(oa, host(out_b))
    want = np.packbits((oa > 0).reshape(-1, 8), axis=1, bitorder="little").reshape(-1)
    assert np.array_equal(host(bits), want)
    res = []
    for use_bits in (False, True):
        dsums = torch.zeros(2 * Cc * 8, dtype=torch.float64, device="cuda:0")
        coef = torch.zeros(2 * Cc, device="cuda:0")
        dgamma, dbeta = torch.zeros(Cc, device="cuda:0"), torch.zeros(Cc, device="cuda:0")
        dxa = empty_act(shape, tdt, fill=5.0)
        dra = act(dres0, tdt)
        if use_bits:
            call("basi_bn_bwd_reduce_bits", da.ref, bits.data_ptr(), xa.ref, bnp.data_ptr(), dsums.data_ptr(),
                 C.c_double(R), dgamma.data_ptr(), dbeta.data_ptr(), coef.data_ptr(), cnt.data_ptr() + 8)
            call("basi_bn_bwd_apply_bits", da.ref, bits.data_ptr(), xa.ref, bnp.data_ptr(), coef.data_ptr(), dxa.ref,
                 dra.ref, 1)
        else:
            call("basi_bn_bwd_reduce", da.ref, out_a.ref, xa.ref, bnp.data_ptr(), 0, dsums.data_ptr(), C.c_double(R),
                 dgamma.data_ptr(), dbeta.data_ptr(), coef.data_ptr(), cnt.data_ptr() + 12)
            call("basi_bn_bwd_apply", da.ref, out_a.ref, xa.ref, bnp.data_ptr(), coef.data_ptr(), 0, dxa.ref, dra.ref, 1)
        res.append((host(dxa), host(dra), host(dgamma), host(dbeta)))
    assert rel_err(res[1][0], res[0][0]) < 1e-2           # same arithmetic; only the atomics order differs
    assert np.array_equal(res[1][1], res[0][1])           # dres += dout * mask: exact
    assert rel_err(res[1][2], res[0][2]) < 1e-5 and rel_err(res[1][3], res[0][3]) < 1e-5


# ------------------------------------------------------------------ pyramid pooling in one pass
@pytest.mark.parametrize("dtype", ["f32", "bf16"])
@pytest.mark.parametrize("B,H,Cc,ks", [(2, 40, 128, (40, 20, 13, 6)), (3, 8, 64, (8, 4, 2, 1)), (1, 23, 16, (23, 11, 7)),
                                       (16, 40, 1024, (40, 20, 13, 6))])
def test_avgpool_multi_forward_backward(dtype, B, H, Cc, ks):
    from gpu_util import act, bf16_round, call, empty_act, host, rel_err
    from basi_b200 import _lib
    from basi_b200._lib import Tensor
    td = torch.float32 if dtype == "f32" else torch.bfloat16
    rng = np.random.RandomState(17)
    x = _u(rng, B, H, H, Cc)
    if dtype == "bf16":
        x = bf16_round(x)
    xa = act(x, td)
    ys = [empty_act((B, H // k, H // k, Cc), td, fill=9.0) for k in ks]
    karr = (C.c_int * len(ks))(*ks)
    yptr = (C.POINTER(Tensor) * len(ks))(*[C.pointer(y.desc) for y in ys])
    nfl = _lib.load().basi_avgpool_multi_scratch_floats(xa.ref, len(ks), karr)
    assert nfl == B * H * sum(H // k for k in ks) * Cc         # per-row window sums [n][h][row cells][c]
    scratch = torch.full((nfl,), 5.0, dtype=torch.float32, device="cuda:0")   # (needs no initialisation)
    tol = 1e-5 if dtype == "f32" else 1e-2
    first = None
    for _ in range(3):                                      # no atomics: bit-identical from call to call
        call("basi_avgpool_multi_fwd", xa.ref, len(ks), karr, yptr, scratch.data_ptr())
        for k, y in zip(ks, ys):
            want = nhwc(O.avg_pool(nchw(x).double(), k))
            assert rel_err(host(y), want) < tol
        got = [host(y).copy() for y in ys]
        if first is None:
            first = got
        assert all(np.array_equal(a, b) for a, b in zip(first, got))
    # adjoint: every pool adds its share to dx in one pass
    dys = [_u(rng, B, H // k, H // k, Cc) for k in ks]
    if dtype == "bf16":
        dys = [bf16_round(d) for d in dys]
    dya = [act(d, td) for d in dys]
    dptr = (C.POINTER(Tensor) * len(ks))(*[C.pointer(d.desc) for d in dya])
    want = np.zeros((B, H, H, Cc), np.float64)
    for k, d in zip(ks, dys):
        o = H // k
        up = np.repeat(np.repeat(d.astype(np.float64), k, axis=1), k, axis=2) / (k * k)
        want[:, :o * k, :o * k, :] += up
    dxa = empty_act((B, H, H, Cc), td, fill=3.0)
    call("basi_avgpool_multi_bwd", dptr, len(ks), karr, dxa.ref, 0)
    assert rel_err(host(dxa), want) < tol
    base = _u(rng, B, H, H, Cc)
    if dtype == "bf16":
        base = bf16_round(base)
    dxb = act(base, td)
    call("basi_avgpool_multi_bwd", dptr, len(ks), karr, dxb.ref, 1)
    assert rel_err(host(dxb), want + base) < 2 * tol


# ------------------------------------------------------------------ cooperative streamed BN backward (one launch)
@pytest.mark.parametrize("dtype", ["f32", "bf16"])
@pytest.mark.parametrize("mask_mode,dres_mode", [("bits", "acc"), ("bits", "none"), ("out", "write"), ("out", "acc"),
                                                 ("from_x", "none"), ("none", "none")])
@pytest.mark.parametrize("shape", [(16, 40, 40, 512), (3, 37, 41, 128)])
def test_batch_norm_backward_coop(dtype, mask_mode, dres_mode, shape):
    """basi_bn_bwd_coop == basi_bn_bwd_reduce[_bits] + basi_bn_bwd_apply[_bits] for every mask / residual mode."""
    from gpu_util import act, bf16_round, call, dev, empty_act, host, rel_err
    from basi_b200 import _lib
    if mask_mode == "bits" and dtype == "f32":
        pytest.skip("mask bits are bf16 only")
    B, H, W, Cc = shape
    rng = np.random.RandomState(23)
    tdt = torch.float32 if dtype == "f32" else torch.bfloat16
    rnd = (lambda a: a) if dtype == "f32" else bf16_round
    x, x2 = rnd(_u(rng, *shape) * 2 + 0.5), rnd(_u(rng, *shape))
    dout, dres0 = rnd(_u(rng, *shape)), rnd(_u(rng, *shape))
    gamma, beta = rng.uniform(0.5, 1.5, Cc).astype(np.float32), _u(rng, Cc)
    R = float(B * H * W)
    xa, x2a, da = act(x, tdt), act(x2, tdt), act(dout, tdt)
    has_out, has_bits = int(mask_mode == "out"), int(mask_mode == "bits")
    dacc = 1 if dres_mode == "acc" else 0
    if _lib.load().basi_bn_bwd_coop_supported(xa.ref, has_out, has_bits, dacc) != 1:
        pytest.skip("tensor does not take the cooperative path on this device")
    gd, bd = dev(gamma), dev(beta)
    sums = torch.zeros(2 * Cc * 8, dtype=torch.float64, device="cuda:0")
    bnp = torch.zeros(4 * Cc, device="cuda:0")
    cnt = torch.zeros(24, dtype=torch.int32, device="cuda:0")
    call("basi_bn_stats", xa.ref, sums.data_ptr(), gd.data_ptr(), bd.data_ptr(), C.c_double(R), C.c_float(1e-5),
         bnp.data_ptr(), cnt.data_ptr())
    outa = empty_act(shape, tdt)
    bits = torch.zeros(max(1, B * H * W * Cc // 8), dtype=torch.uint8, device="cuda:0")
    junction = mask_mode in ("bits", "out")
    if mask_mode == "bits":
        call("basi_bn_apply_bits", xa.ref, bnp.data_ptr(), x2a.ref, None, 1, outa.ref, bits.data_ptr())
    elif junction:
        call("basi_bn_apply", xa.ref, bnp.data_ptr(), x2a.ref, None, 1, outa.ref)
    from_x = 1 if mask_mode == "from_x" else 0
    res = []
    for coop in (False, True, True):
        dsums = torch.zeros(2 * Cc * 8, dtype=torch.float64, device="cuda:0")
        coef = torch.zeros(2 * Cc, device="cuda:0")
        dgamma, dbeta = torch.full((Cc,), 0.25, device="cuda:0"), torch.full((Cc,), -0.5, device="cuda:0")
        dxa = empty_act(shape, tdt, fill=5.0)
        dra = act(dres0, tdt) if dres_mode != "none" else None
        dres_ref = dra.ref if dra is not None else None
        out_ref = outa.ref if mask_mode == "out" else None
        bits_ptr = bits.data_ptr() if mask_mode == "bits" else None
        if coop:
            call("basi_bn_bwd_coop", da.ref, out_ref, bits_ptr, xa.ref, bnp.data_ptr(), from_x, dsums.data_ptr(),
                 C.c_double(R), dgamma.data_ptr(), dbeta.data_ptr(), coef.data_ptr(), cnt.data_ptr() + 16, dxa.ref,
                 dres_ref, dacc)
        elif mask_mode == "bits":
            call("basi_bn_bwd_reduce_bits", da.ref, bits_ptr, xa.ref, bnp.data_ptr(), dsums.data_ptr(), C.c_double(R),
                 dgamma.data_ptr(), dbeta.data_ptr(), coef.data_ptr(), cnt.data_ptr() + 8)
            call("basi_bn_bwd_apply_bits", da.ref, bits_ptr, xa.ref, bnp.data_ptr(), coef.data_ptr(), dxa.ref, dres_ref,
                 dacc)
        else:
            call("basi_bn_bwd_reduce", da.ref, out_ref, xa.ref, bnp.data_ptr(), from_x, dsums.data_ptr(), C.c_double(R),
                 dgamma.data_ptr(), dbeta.data_ptr(), coef.data_ptr(), cnt.data_ptr() + 12)
            call("basi_bn_bwd_apply", da.ref, out_ref, xa.ref, bnp.data_ptr(), coef.data_ptr(), from_x, dxa.ref,
                 dres_ref, dacc)
        res.append((host(dxa), host(dra) if dra is not None else None, host(dgamma), host(dbeta), host(coef)))
    ptol = 1e-5 if dtype == "f32" else 1e-2
    for k in (1, 2):
        assert rel_err(res[k][0], res[0][0]) < ptol
        if res[0][1] is not None:
            assert np.array_equal(res[k][1], res[0][1])
        for j in (2, 3, 4):
            assert rel_err(res[k][j], res[0][j]) < 1e-5


@pytest.mark.parametrize("out_dtype", ["f32", "bf16"])
@pytest.mark.parametrize("cout,H,W,B", [(32, 64, 64, 2), (64, 33, 47, 1), (32, 320, 320, 2)])
def test_stem_conv_fused_bn_statistics(out_dtype, cout, H, W, B):
    """basi_stem_fprop_stats: same output as basi_conv_fprop, and [mean | istd | gamma*istd | beta] equal to the
    statistics of the tensor as stored."""
    from gpu_util import act, call, dev, empty_act, host, rel_err
    from basi_b200 import _lib
    from basi_b200._lib import ConvDesc
    rng = np.random.RandomState(cout + W)
    x = _u(rng, B, H, W, 4) + 0.2
    w = (_u(rng, 3, 3, 4, cout) / 6.0).astype(np.float32)
    gamma, beta = rng.uniform(0.5, 1.5, cout).astype(np.float32), _u(rng, cout)
    oh, ow = (H + 1) // 2, (W + 1) // 2
    pt, pl = O.tf_same_pad(H, 3, 2, 1)[0], O.tf_same_pad(W, 3, 2, 1)[0]
    desc = ConvDesc(3, 3, 2, 1, pt, pl, 0)
    tdt = torch.float32 if out_dtype == "f32" else torch.bfloat16
    xa = act(x)
    ya, yb = empty_act((B, oh, ow, cout), tdt, fill=7.0), empty_act((B, oh, ow, cout), tdt, fill=7.0)
    assert _lib.load().basi_stem_fprop_stats_supported(C.byref(desc), xa.ref, ya.ref) == 1
    wd, gd, bd = dev(w), dev(gamma), dev(beta)
    sums = torch.zeros(2 * cout * 8, dtype=torch.float64, device="cuda:0")
    bnp = torch.zeros(4 * cout, device="cuda:0")
    cnt = torch.zeros(2, dtype=torch.int32, device="cuda:0")
    call("basi_conv_fprop", C.byref(desc), xa.ref, wd.data_ptr(), None, ya.ref)
    call("basi_stem_fprop_stats", C.byref(desc), xa.ref, wd.data_ptr(), yb.ref, sums.data_ptr(), gd.data_ptr(),
         bd.data_ptr(), C.c_double(B * oh * ow), C.c_float(1e-5), bnp.data_ptr(), cnt.data_ptr())
    y = host(yb)
    assert np.array_equal(host(ya), y)
    flat = y.reshape(-1, cout).astype(np.float64)
    mean, var = flat.mean(0), flat.var(0)
    got = host(bnp).reshape(4, cout)
    istd = 1.0 / np.sqrt(var + 1e-5)
    assert rel_err(got[0], mean) < 1e-5 and rel_err(got[1], istd) < 1e-5
    assert rel_err(got[2], gamma * istd) < 1e-5 and np.array_equal(got[3], beta)


# ---------------------------------------------------------------------------------------------------------------
# A3 / F3: label encodings and click sampling on the device (bit-exact integer work)
# ---------------------------------------------------------------------------------------------------------------
def _instance_maps(B, P, rng, ids=(1, 2, 7, 85, 86, 170, 200, 254)):
    ann = np.zeros((B, P, P), dtype=np.uint8)
    nums = []
    for b in range(B):
        chosen = rng.choice(ids, size=3, replace=False)
        for j, k in enumerate(chosen):
            y0, x0 = rng.randint(0, P - 6), rng.randint(0, P - 6)
            ann[b, y0:y0 + rng.randint(3, 6), x0:x0 + rng.randint(3, 6)] = k
        ann[b, rng.randint(0, P, 20), rng.randint(0, P, 20)] = 255          # border pixels
        present = [k for k in chosen if np.any(ann[b] == k)]
        nums.append(int(present[rng.randint(0, len(present))]))
    return ann, np.asarray(nums, dtype=np.int32)


@pytest.mark.parametrize("mode", ["binary", "border", "three", "coco"])
def test_label_encode_bit_exact(mode):
    from gpu_util import call, dev, host
    rng = np.random.RandomState(11)
    B, P = 5, 40
    ann, nums = _instance_maps(B, P, rng)
    att = None
    if mode == "coco":
        att = (ann == nums[:, None, None]).astype(np.uint8)
        ann = ((ann > 0) & (ann < 255)).astype(np.uint8) * rng.randint(1, 4, size=ann.shape).astype(np.uint8)
        ref = np.stack([O.encode_labels_coco(ann[b], att[b]) for b in range(B)])
    else:
        fn = {"binary": O.encode_labels_binary, "border": O.encode_labels_border, "three": O.encode_labels_three}[mode]
        ref = np.stack([fn(ann[b], nums[b]) for b in range(B)])
    if mode == "border":
        # the uint8 wrap-around and the aliasing quirk of instance ids >= 85 are part of the reference semantics
        assert set(np.unique(ref)) <= {0, 1, 2, 3} and np.all(ref[ann == 0] == 3) and np.all(ref[ann == 255] == 2)
    code = {"binary": 0, "border": 1, "three": 2, "coco": 3}[mode]
    oi = torch.full((B, P, P), -7, dtype=torch.int32, device="cuda:0")
    of = torch.full((B, P, P), -7.0, dtype=torch.float32, device="cuda:0")
    annd, numd = dev(ann), dev(nums)
    attd = dev(att) if att is not None else None
    call("basi_label_encode", annd.data_ptr(), attd.data_ptr() if attd is not None else None,
         None if attd is not None else numd.data_ptr(), code, oi.data_ptr(), of.data_ptr(), B, C.c_int64(P * P))
    assert np.array_equal(host(oi), ref.astype(np.int32))
    assert np.array_equal(host(of), ref.astype(np.float32))


@pytest.mark.parametrize("f32", [0, 1])
def test_click_sampling_matches_argwhere(f32):
    """k-th pixel (row-major) of the attended instance times the ratio == np.argwhere(ann == 1)[k] * ratio, for
    every valid k of small maps and for the RNG-drawn k of the reference's sampling loop."""
    from gpu_util import call, dev, host
    rng = np.random.RandomState(3)
    B, H, W, ratio = 6, 23, 40, 8
    lab = (rng.rand(B, H, W) < 0.07).astype(np.int32)
    lab[0] = 0
    lab[0, H - 1, W - 1] = 1                     # a single pixel, the last one
    lab[1] = 1                                   # every pixel
    labd = dev(lab.astype(np.float32) if f32 else lab)
    counts = torch.zeros(B, dtype=torch.int32, device="cuda:0")
    call("basi_click_count", labd.data_ptr(), f32, 1, B, H * W, counts.data_ptr())
    cnt = host(counts)
    assert np.array_equal(cnt, lab.reshape(B, -1).sum(1))
    clicks = torch.zeros((B, 2), dtype=torch.int32, device="cuda:0")
    for trial in range(6):
        k = np.asarray([0 if trial == 0 else (c - 1 if trial == 1 else rng.randint(0, c)) for c in cnt], dtype=np.int32)
        kd = dev(k)
        call("basi_click_select", labd.data_ptr(), f32, 1, B, H, W, kd.data_ptr(), ratio, clicks.data_ptr())
        got = host(clicks)
        ref = np.asarray([O.sample_click(lab[b], int(k[b]), ratio) for b in range(B)], dtype=np.int32)
        assert np.array_equal(got, ref), (trial, got, ref)
    kd = dev(np.asarray(cnt, dtype=np.int32))    # k == count: out of range -> (-1, -1)
    call("basi_click_select", labd.data_ptr(), f32, 1, B, H, W, kd.data_ptr(), ratio, clicks.data_ptr())
    assert np.all(host(clicks) == -1)


def test_engine_label_input_feeds_border_labels_and_reference_clicks():
    """Engine.feed_annotations: uint8 instance maps -> 4-class labels + clicks on the device, clicks identical to the
    reference's host loop run with the same numpy RNG state."""
    from basi_b200.BAISPSPNet import PSPNet, Placeholder
    from basi_b200.engine import Engine
    S, F, B, P = 64, 8, 3, 8
    net = PSPNet({'data': Placeholder((None, S, S, 4))}, num_classes=21, num_segment=4, is_training=True,
                 last_pool_size=P, filter_number=F, variant="4BorderClass")
    eng = Engine(net, B, "bf16", True, dict(kind="softmax", class_weight=0.1))
    eng.enable_click_input(30)
    eng.enable_label_input("border")
    rng = np.random.RandomState(5)
    ann, nums = _instance_maps(B, P, rng, ids=(1, 2, 3))
    k = eng.feed_annotations(ann, nums, rng=np.random.RandomState(42))
    torch.cuda.synchronize()
    ref_lab = np.stack([O.encode_labels_border(ann[b], nums[b]) for b in range(B)])
    assert np.array_equal(eng.label_seg.cpu().numpy()[..., 0], ref_lab.astype(np.int32))
    r2 = np.random.RandomState(42)
    ref_clicks = []
    for b in range(B):                                   # the reference loop (BAISData.py:63-66)
        where = np.argwhere(ref_lab[b] == 1)
        where = where[r2.randint(0, len(where))]
        ref_clicks.append([where[0] * 8, where[1] * 8])
    assert np.array_equal(eng.clicks_dev.cpu().numpy(), np.asarray(ref_clicks, dtype=np.int32))


# ------------------------------------------------------------------ F4 ops (cascade): sigmoid, channel-select weighted BCE
def _guarded(n, dtype=torch.float32, guard=64, fill=-77.0):
    """a device buffer of n elements between two guard bands (the kernels must not touch the bands)"""
    full = torch.full((n + 2 * guard,), fill, dtype=dtype, device="cuda:0")
    return full, full[guard:guard + n]


def _guards_intact(full, n, guard=64, fill=-77.0):
    torch.cuda.synchronize()
    return bool((full[:guard] == fill).all()) and bool((full[guard + n:] == fill).all())


@pytest.mark.parametrize("n", [1, 7, 4 * 40 * 40 * 4, 100003])
def test_sigmoid_forward_backward(n):
    from gpu_util import call, dev, host
    rng = np.random.RandomState(n)
    x = (rng.randn(n) * 6).astype(np.float32)
    x[: min(n, 4)] = [0.0, -90.0, 90.0, 1e-8][: min(n, 4)]
    dy = rng.randn(n).astype(np.float32)
    xt = torch.from_numpy(x).double().requires_grad_(True)
    yt = torch.sigmoid(xt)
    (yt * torch.from_numpy(dy).double()).sum().backward()
    xd, dyd = dev(x), dev(dy)
    fy, y = _guarded(n)
    call("basi_sigmoid_fwd", xd.data_ptr(), y.data_ptr(), C.c_int64(n))
    assert _guards_intact(fy, n)
    assert np.max(np.abs(host(y) - yt.detach().numpy())) < 2e-7
    fdx, dx = _guarded(n)
    call("basi_sigmoid_bwd", dyd.data_ptr(), y.data_ptr(), dx.data_ptr(), C.c_int64(n), 0)
    assert _guards_intact(fdx, n)
    assert np.max(np.abs(host(dx) - xt.grad.numpy())) < 1e-6
    call("basi_sigmoid_bwd", dyd.data_ptr(), y.data_ptr(), dx.data_ptr(), C.c_int64(n), 1)      # accumulate
    assert np.max(np.abs(host(dx) - 2 * xt.grad.numpy())) < 2e-6
    assert _guards_intact(fdx, n)


@pytest.mark.parametrize("rows,Cn,sel", [(1, 2, 1), (4 * 40 * 40, 2, 1), (1001, 4, 2), (50000, 2, 0)])
def test_wbce_channel_select(rows, Cn, sel):
    """cal_loss of the cascade: weighted_cross_entropy_with_logits on tf.split(segment, C, axis=3)[sel]
    (back/8AttentionU/BAISRunnerTrain.py:176-180); the gradient goes to channel `sel`, zeros to the others."""
    from gpu_util import call, dev, host
    rng = np.random.RandomState(rows)
    x = rng.rand(rows, Cn).astype(np.float32)                  # (sigmoid outputs are what the reference feeds)
    z = (rng.rand(rows) < 0.3).astype(np.float32)
    pw, lw, gw = 3.0, 2.0 / (4 * rows), 0.37
    xt = torch.from_numpy(x).double().requires_grad_(True)
    loss = O.weighted_cross_entropy_with_logits(torch.from_numpy(z).double(), xt[:, sel], pw).sum() * lw
    loss.backward()
    xd, zd = dev(x), dev(z)
    acc = torch.zeros(4, dtype=torch.float64, device="cuda:0")
    fg, g = _guarded(rows * Cn)
    call("basi_wbce_sel_fwd_bwd", xd.data_ptr(), Cn, sel, zd.data_ptr(), C.c_float(pw), C.c_double(lw), C.c_float(gw),
         C.c_int64(rows), acc.data_ptr(), g.data_ptr())
    assert _guards_intact(fg, rows * Cn)
    got = host(g).reshape(rows, Cn)
    assert abs(float(acc[0]) - float(loss.detach())) < 1e-6 * max(1.0, abs(float(loss.detach())))
    want = xt.grad.numpy() * (gw / lw)
    assert np.max(np.abs(got - want)) < 1e-5 * max(1.0, np.abs(want).max())
    assert np.all(got[:, [c for c in range(Cn) if c != sel]] == 0.0)


def test_deterministic_pool_and_gemv_do_not_touch_their_neighbours():
    """The rows/cells pyramid pooling writes exactly basi_avgpool_multi_scratch_floats() floats of scratch and the
    workspace GEMV exactly basi_skinny_fwd_workspace_floats() (guard bands around both stay intact)."""
    from gpu_util import act, call, dev, empty_act, host
    from basi_b200 import _lib
    from basi_b200._lib import Tensor
    rng = np.random.RandomState(2)
    B, H, Cc, ks = 3, 40, 64, (40, 20, 13, 6)
    xa = act(_u(rng, B, H, H, Cc), torch.bfloat16)
    ys = [empty_act((B, H // k, H // k, Cc), torch.bfloat16, fill=9.0) for k in ks]
    karr = (C.c_int * len(ks))(*ks)
    yptr = (C.POINTER(Tensor) * len(ks))(*[C.pointer(y.desc) for y in ys])
    nfl = int(_lib.load().basi_avgpool_multi_scratch_floats(xa.ref, len(ks), karr))
    full, scratch = _guarded(nfl)
    call("basi_avgpool_multi_fwd", xa.ref, len(ks), karr, yptr, scratch.data_ptr())
    assert _guards_intact(full, nfl)
    M, K, N = 5, 25 * 64, 96
    a, w, b = _u(rng, M, K), _u(rng, K, N) / 40, _u(rng, N)
    nws = int(_lib.load().basi_skinny_fwd_workspace_floats(M, K, N))
    fullw, ws = _guarded(nws)
    fully, y = _guarded(M * N)
    ad, wd, bd = dev(a), dev(w), dev(b)
    call("basi_skinny_fwd_ws", ad.data_ptr(), 0, C.c_int64(K), wd.data_ptr(), bd.data_ptr(), y.data_ptr(), M, K, N, 0,
         ws.data_ptr())
    assert _guards_intact(fullw, nws) and _guards_intact(fully, M * N)
    assert np.max(np.abs(host(y).reshape(M, N) - (a.astype(np.float64) @ w + b))) < 1e-4
