"""GPU parity of the tcgen05/TMEM/TMA implicit-GEMM convolutions (fprop, dgrad, wgrad) against the oracle."""
import ctypes as C

import numpy as np
import pytest
import torch

from oracle import basi_oracle as O

pytestmark = pytest.mark.gpu

# k, dil, Cin, Cout, H, W, B
TC_CASES = [
    (1, 1, 64, 128, 16, 16, 2),
    (1, 1, 64, 32, 80, 80, 1),      # conv2 reduce, N tile 32
    (3, 1, 64, 64, 40, 40, 2),      # conv3 3x3 (tile 8x8x2)
    (3, 2, 128, 128, 40, 40, 3),    # conv4 dilated, odd batch -> image-axis overhang
    (3, 4, 256, 256, 24, 24, 1),    # conv5 dilated, B=1
    (1, 1, 512, 128, 40, 40, 2),
    (1, 1, 128, 512, 40, 40, 2),
    (3, 1, 256, 64, 20, 12, 2),     # ragged map -> partial tiles
    (3, 1, 32, 32, 80, 80, 1),      # conv2 3x3: 32-channel K/N -> zero-filled 64-wide boxes
    (1, 1, 32, 128, 80, 80, 2),     # conv2 increase
    (3, 1, 32, 64, 48, 48, 1),      # conv1_3-like
    (3, 1, 64, 96, 16, 16, 2),      # Cout not a multiple of 64/128 -> partial co tile in wgrad
    (3, 4, 256, 256, 40, 40, 6),    # conv5 3x3 at 75 pixel tiles: 256-wide tiles, 36 k-steps -> CTA pairs in auto mode,
                                    # odd tile count (the last pair has a phantom second tile)
]


def _tc_case(case, sliced=False):
    from basi_b200 import _lib
    from basi_b200._lib import ConvDesc
    from basi_b200.engine import Act
    from gpu_util import bf16_round, call, dev, host
    k, d, cin, cout, H, W, B = case
    rng = np.random.RandomState(sum(case))
    x = bf16_round(rng.uniform(-1, 1, (B, H, W, cin)).astype(np.float32))
    w = bf16_round((rng.uniform(-1, 1, (k, k, cin, cout)) / np.sqrt(k * k * cin)).astype(np.float32))
    dy = bf16_round(rng.uniform(-1, 1, (B, H, W, cout)).astype(np.float32))
    pad = d * (k - 1) // 2
    desc = ConvDesc(k, k, 1, d, pad, pad, 0)
    bt = torch.bfloat16
    if sliced:   # operands are channel slices of wider buffers (the PSP concat case)
        xw = torch.zeros((B, H, W, cin + 128), dtype=bt, device="cuda:0")
        xw[..., 64:64 + cin] = torch.from_numpy(x).to("cuda:0").to(bt)
        xa = Act(xw[..., 64:64 + cin])
        yw = torch.full((B, H, W, cout + 64), 3.0, dtype=bt, device="cuda:0")
        ya = Act(yw[..., 32:32 + cout])
    else:
        xa = Act(torch.from_numpy(x).to("cuda:0").to(bt).contiguous())
        ya = Act(torch.full((B, H, W, cout), 3.0, dtype=bt, device="cuda:0"))
    wd = dev(w)
    w_io = torch.zeros(k * k * cin * cout, dtype=bt, device="cuda:0")
    w_oi = torch.zeros(k * k * cin * cout, dtype=bt, device="cuda:0")
    call("basi_tc_pack_weights", wd.data_ptr(), w_io.data_ptr(), w_oi.data_ptr(), k * k, cin, cout)
    lib = _lib.load()
    res = {}
    xt = torch.from_numpy(x).permute(0, 3, 1, 2).double().requires_grad_(True)
    wt = torch.from_numpy(w).double().requires_grad_(True)
    yref = O.conv2d(xt, wt, 1, pad, d)
    (yref * torch.from_numpy(dy).permute(0, 3, 1, 2).double()).sum().backward()
    res["y_ref"] = yref.detach().permute(0, 2, 3, 1).numpy()
    res["dx_ref"] = xt.grad.permute(0, 2, 3, 1).numpy()
    res["dw_ref"] = wt.grad.numpy()
    handles = []
    if lib.basi_tc_conv_supported(0, C.byref(desc), xa.ref, ya.ref) == 1:
        h = C.c_void_p()
        _lib.call("basi_tc_conv_create", 0, C.byref(desc), xa.ref, ya.ref, w_oi.data_ptr(), None, 0, C.byref(h))
        call("basi_tc_conv_run", h)
        res["y"] = host(ya)
        if sliced:
            full = host(yw)
            assert np.all(full[..., :32] == 3.0) and np.all(full[..., 32 + cout:] == 3.0)
        handles.append(h)
    dya = Act(torch.from_numpy(dy).to("cuda:0").to(bt).contiguous())
    dxa = Act(torch.full((B, H, W, cin), 1.0, dtype=bt, device="cuda:0"))
    if lib.basi_tc_conv_supported(1, C.byref(desc), xa.ref, ya.ref) == 1:
        h = C.c_void_p()
        _lib.call("basi_tc_conv_create", 1, C.byref(desc), dya.ref, dxa.ref, w_io.data_ptr(), None, 0, C.byref(h))
        call("basi_tc_conv_run", h)
        res["dx"] = host(dxa)
        h2 = C.c_void_p()
        _lib.call("basi_tc_conv_create", 1, C.byref(desc), dya.ref, dxa.ref, w_io.data_ptr(), None, 1, C.byref(h2))
        call("basi_tc_conv_run", h2)
        res["dx_acc"] = host(dxa)
        handles += [h, h2]
    if lib.basi_tc_conv_supported(2, C.byref(desc), xa.ref, ya.ref) == 1:
        dw = torch.zeros((k, k, cin, cout), dtype=torch.float32, device="cuda:0")
        h = C.c_void_p()
        _lib.call("basi_tc_conv_create", 2, C.byref(desc), xa.ref, dya.ref, None, dw.data_ptr(), 1, C.byref(h))
        call("basi_tc_conv_run", h)
        res["dw"] = host(dw)
        handles.append(h)
    torch.cuda.synchronize()
    for h in handles:
        lib.basi_tc_conv_destroy(h)
    return res


# tile scheduling variants of the fprop/dgrad kernel, selected at plan creation through the environment
TC_MODES = {
    "auto": {},
    "single": {"BASI_TC_MT": "1", "BASI_TC_CLUSTER": "0"},     # one 128-pixel tile per work item
    "double": {"BASI_TC_MT": "2", "BASI_TC_CLUSTER": "0"},     # two pixel tiles share the weight box
    "pairs": {"BASI_TC_CLUSTER": "1"},                          # cta_group::2: one M = 256 MMA per CTA pair
    "halo": {"BASI_TC_HALO": "1", "BASI_TC_CLUSTER": "0"},      # 3x3: one activation halo box serves all nine taps
}


@pytest.mark.parametrize("mode", list(TC_MODES))
@pytest.mark.parametrize("case", TC_CASES, ids=lambda c: "k%dd%d_%dto%d_%dx%d_B%d" % c)
def test_tc_conv_matches_oracle(case, mode, monkeypatch):
    from gpu_util import rel_err
    monkeypatch.setenv("BASI_EXPERIMENTS", "1")       # scheduling switches are experiment-gated
    for k_, v_ in TC_MODES[mode].items():
        monkeypatch.setenv(k_, v_)
    r = _tc_case(case)
    k, d, cin, cout = case[:4]
    assert ("y" in r) == (cin % 8 == 0 and cout % 32 == 0), "fprop support does not match the documented rule"
    assert ("dx" in r) == (cout % 8 == 0 and cin % 32 == 0), "dgrad support does not match the documented rule"
    assert "dw" in r, "wgrad support does not match the documented rule"
    assert rel_err(r["y"], r["y_ref"]) < 1e-2             # bf16 output rounding only (fp32 accumulate)
    if "dx" in r:
        assert rel_err(r["dx"], r["dx_ref"]) < 1e-2
        assert rel_err(r["dx_acc"], 2 * r["dx_ref"]) < 2e-2
    if "dw" in r:
        assert rel_err(r["dw"], r["dw_ref"]) < 1e-4      # fp32 accumulate + fp32 atomics, bf16-exact inputs


def test_tc_conv_on_channel_slices():
    from gpu_util import rel_err
    r = _tc_case((3, 1, 128, 64, 16, 16, 2), sliced=True)
    assert rel_err(r["y"], r["y_ref"]) < 1e-2
    assert rel_err(r["dw"], r["dw_ref"]) < 1e-4


def test_tc_refuses_unsupported_geometries():
    from basi_b200 import _lib
    from basi_b200._lib import ConvDesc
    from basi_b200.engine import Act
    lib = _lib.load()
    bt = torch.bfloat16
    x = Act(torch.zeros((1, 16, 16, 64), dtype=bt, device="cuda:0"))
    y = Act(torch.zeros((1, 8, 8, 64), dtype=bt, device="cuda:0"))
    assert lib.basi_tc_conv_supported(0, C.byref(ConvDesc(1, 1, 2, 1, 0, 0, 0)), x.ref, y.ref) == 0     # stride 2
    x4 = Act(torch.zeros((1, 16, 16, 32), dtype=bt, device="cuda:0"))
    y4 = Act(torch.zeros((1, 16, 16, 24), dtype=bt, device="cuda:0"))
    assert lib.basi_tc_conv_supported(0, C.byref(ConvDesc(3, 3, 1, 1, 1, 1, 0)), x4.ref, y4.ref) == 0   # Cout 24
    h = C.c_void_p()
    rc = lib.basi_tc_conv_create(0, C.byref(ConvDesc(3, 3, 1, 1, 1, 1, 0)), x4.ref, y4.ref, None, None, 0, C.byref(h))
    assert rc == -1 and b"not supported" in lib.basi_last_error()


@pytest.mark.parametrize("mode", list(TC_MODES))
@pytest.mark.parametrize("case", [(3, 2, 128, 128, 40, 40, 3), (1, 1, 128, 512, 40, 40, 2), (1, 1, 64, 32, 80, 80, 1)],
                         ids=lambda c: "k%dd%d_%dto%d_%dx%d_B%d" % c)
def test_tc_fprop_fused_bn_statistics(case, mode, monkeypatch):
    """The fprop epilogue accumulates per-channel sum / sum-of-squares of the stored (bf16) output and the last CTA
    finalizes [mean | istd | gamma*istd | beta]: must equal statistics computed from the stored tensor."""
    from basi_b200 import _lib
    from basi_b200._lib import ConvDesc
    from basi_b200.engine import Act
    from gpu_util import bf16_round, call, dev, host, rel_err
    monkeypatch.setenv("BASI_EXPERIMENTS", "1")       # scheduling switches are experiment-gated
    for k_, v_ in TC_MODES[mode].items():
        monkeypatch.setenv(k_, v_)
    k, d, cin, cout, H, W, B = case
    rng = np.random.RandomState(3)
    x = bf16_round(rng.uniform(-1, 1, (B, H, W, cin)).astype(np.float32) + 0.3)
    w = bf16_round((rng.uniform(-1, 1, (k, k, cin, cout)) / np.sqrt(k * k * cin)).astype(np.float32))
    gamma, beta = rng.uniform(0.5, 1.5, cout).astype(np.float32), rng.uniform(-1, 1, cout).astype(np.float32)
    pad = d * (k - 1) // 2
    desc = ConvDesc(k, k, 1, d, pad, pad, 0)
    bt = torch.bfloat16
    xa = Act(torch.from_numpy(x).to("cuda:0").to(bt).contiguous())
    ya = Act(torch.zeros((B, H, W, cout), dtype=bt, device="cuda:0"))
    wd, gd, bd = dev(w), dev(gamma), dev(beta)
    w_io = torch.zeros(k * k * cin * cout, dtype=bt, device="cuda:0")
    w_oi = torch.zeros(k * k * cin * cout, dtype=bt, device="cuda:0")
    call("basi_tc_pack_weights", wd.data_ptr(), w_io.data_ptr(), w_oi.data_ptr(), k * k, cin, cout)
    sums = torch.zeros(2 * cout * 8, dtype=torch.float64, device="cuda:0")
    bnp = torch.zeros(4 * cout, device="cuda:0")
    cnt = torch.zeros(2, dtype=torch.int32, device="cuda:0")
    h = C.c_void_p()
    _lib.call("basi_tc_conv_create", 0, C.byref(desc), xa.ref, ya.ref, w_oi.data_ptr(), None, 0, C.byref(h))
    R = float(B * H * W)
    _lib.call("basi_tc_conv_set_bn_stats", h, sums.data_ptr(), gd.data_ptr(), bd.data_ptr(), C.c_double(R),
              C.c_float(1e-5), bnp.data_ptr(), cnt.data_ptr())
    call("basi_tc_conv_run", h)
    y = host(ya).astype(np.float64).reshape(-1, cout)
    got, s = host(bnp), host(sums).reshape(8, 2 * cout).sum(0)
    _lib.load().basi_tc_conv_destroy(h)
    assert rel_err(s[:cout], y.sum(0)) < 1e-5 and rel_err(s[cout:], (y * y).sum(0)) < 1e-5
    mean, var = y.mean(0), y.var(0)
    istd = 1 / np.sqrt(var + 1e-5)
    assert rel_err(got[:cout], mean) < 1e-5
    assert rel_err(got[cout:2 * cout], istd) < 1e-4
    assert rel_err(got[2 * cout:3 * cout], gamma * istd) < 1e-4
    assert np.array_equal(got[3 * cout:], beta)
    assert int(host(cnt)[0]) > 0


# ---------------------------------------------------------------------------------------------------------------
# split-operand ("bf16x6") mode: float32 tensors, bf16 [hi|mid|lo] operands, fp32-grade products on tcgen05
# ---------------------------------------------------------------------------------------------------------------
SPLIT_CASES = [
    (1, 1, 64, 128, 16, 16, 2),
    (3, 1, 32, 32, 80, 80, 1),      # 32-channel layers: [hi|mid] share a 64-channel box
    (3, 2, 128, 128, 40, 40, 3),    # conv4 dilated, odd batch
    (1, 1, 512, 128, 40, 40, 2),
    (1, 1, 128, 512, 40, 40, 2),
    (3, 1, 256, 64, 20, 12, 2),     # ragged map
    (1, 1, 32, 128, 80, 80, 2),
    (3, 4, 256, 256, 24, 24, 1),
]


def _split_case(case, with_stats=False, parts=3):
    from basi_b200 import _lib
    from basi_b200._lib import ConvDesc, PackEntry
    from basi_b200.engine import Act
    from gpu_util import call, dev, host
    k, d, cin, cout, H, W, B = case
    rng = np.random.RandomState(sum(case) + 1)
    x = rng.uniform(-1, 1, (B, H, W, cin)).astype(np.float32)
    w = (rng.uniform(-1, 1, (k, k, cin, cout)) / np.sqrt(k * k * cin)).astype(np.float32)
    dy = rng.uniform(-1, 1, (B, H, W, cout)).astype(np.float32)
    pad = d * (k - 1) // 2
    desc = ConvDesc(k, k, 1, d, pad, pad, 0)
    lib = _lib.load()
    f32, bt = torch.float32, torch.bfloat16
    xa = Act(torch.from_numpy(x).to("cuda:0"))
    ya = Act(torch.full((B, H, W, cout), 3.0, dtype=f32, device="cuda:0"))
    dya = Act(torch.from_numpy(dy).to("cuda:0"))
    dxa = Act(torch.full((B, H, W, cin), 1.0, dtype=f32, device="cuda:0"))
    x3 = Act(torch.zeros((B, H, W, parts * cin), dtype=bt, device="cuda:0"))
    dy3 = Act(torch.zeros((B, H, W, parts * cout), dtype=bt, device="cuda:0"))
    call("basi_split3_bf16", xa.ref, x3.ref)
    call("basi_split3_bf16", dya.ref, dy3.ref)
    # the parts add up to the float32 value (to 2^-24 relative)
    psum = host(x3).astype(np.float64).reshape(B, H, W, parts, cin).sum(3)
    assert np.max(np.abs(psum - x)) <= (2.0 ** -23 if parts == 3 else 2.0 ** -16) * np.max(np.abs(x))
    wd = dev(w)
    kf, kd = lib.basi_tc_split_kcols(cin, parts), lib.basi_tc_split_kcols(cout, parts)
    w_io = torch.zeros(k * k * cin * kd, dtype=bt, device="cuda:0")
    w_oi = torch.zeros(k * k * cout * kf, dtype=bt, device="cuda:0")
    tco, tci = -(-cout // 32), -(-cin // 32)
    ent = (PackEntry * 1)(PackEntry(wd.data_ptr(), w_io.data_ptr(), w_oi.data_ptr(), k * k, cin, cout, 0, tco, tci, 1 if parts == 3 else 2, 0))
    table = torch.from_numpy(np.frombuffer(bytes(ent), dtype=np.uint8).copy()).to("cuda:0")
    call("basi_tc_pack_weights_multi", table.data_ptr(), 1, k * k * tco * tci)
    res = {}
    xt = torch.from_numpy(x).permute(0, 3, 1, 2).double().requires_grad_(True)
    wt = torch.from_numpy(w).double().requires_grad_(True)
    yref = O.conv2d(xt, wt, 1, pad, d)
    (yref * torch.from_numpy(dy).permute(0, 3, 1, 2).double()).sum().backward()
    res["y_ref"] = yref.detach().permute(0, 2, 3, 1).numpy()
    res["dx_ref"] = xt.grad.permute(0, 2, 3, 1).numpy()
    res["dw_ref"] = wt.grad.numpy()
    handles = []
    assert lib.basi_tc_conv_supported_split(0, C.byref(desc), xa.ref, ya.ref) == 1
    h = C.c_void_p()
    _lib.call("basi_tc_conv_create_split", 0, C.byref(desc), x3.ref, ya.ref, w_oi.data_ptr(), None, 0, parts,
              C.byref(h))
    if with_stats:
        gamma, beta = rng.uniform(0.5, 1.5, cout).astype(np.float32), rng.uniform(-1, 1, cout).astype(np.float32)
        gd, bd = dev(gamma), dev(beta)
        sums = torch.zeros(2 * cout * 8, dtype=torch.float64, device="cuda:0")
        bnp = torch.zeros(4 * cout, device="cuda:0")
        cnt = torch.zeros(2, dtype=torch.int32, device="cuda:0")
        _lib.call("basi_tc_conv_set_bn_stats", h, sums.data_ptr(), gd.data_ptr(), bd.data_ptr(),
                  C.c_double(float(B * H * W)), C.c_float(1e-5), bnp.data_ptr(), cnt.data_ptr())
    call("basi_tc_conv_run", h)
    res["y"] = host(ya)
    handles.append(h)
    if with_stats:
        res["bnp"], res["gamma"], res["beta"] = host(bnp), gamma, beta
    if lib.basi_tc_conv_supported_split(1, C.byref(desc), xa.ref, ya.ref) == 1:
        h = C.c_void_p()
        _lib.call("basi_tc_conv_create_split", 1, C.byref(desc), dy3.ref, dxa.ref, w_io.data_ptr(), None, 0, parts,
                  C.byref(h))
        call("basi_tc_conv_run", h)
        res["dx"] = host(dxa)
        h2 = C.c_void_p()
        _lib.call("basi_tc_conv_create_split", 1, C.byref(desc), dy3.ref, dxa.ref, w_io.data_ptr(), None, 1, parts,
                  C.byref(h2))
        call("basi_tc_conv_run", h2)
        res["dx_acc"] = host(dxa)
        handles += [h, h2]
    assert lib.basi_tc_conv_supported_split(2, C.byref(desc), xa.ref, ya.ref) == 1
    dw = torch.zeros((k, k, cin, cout), dtype=f32, device="cuda:0")
    h = C.c_void_p()
    _lib.call("basi_tc_conv_create_split", 2, C.byref(desc), x3.ref, dy3.ref, None, dw.data_ptr(), 1, parts, C.byref(h))
    call("basi_tc_conv_run", h)
    res["dw"] = host(dw)
    handles.append(h)
    torch.cuda.synchronize()
    for h in handles:
        lib.basi_tc_conv_destroy(h)
    return res


@pytest.mark.parametrize("case", SPLIT_CASES[:5], ids=lambda c: "k%dd%d_%dto%d_%dx%d_B%d" % c)
def test_tc_split2_conv_three_products(case):
    """Two-part split ([hi|mid], three products): 16 mantissa bits per operand; the unbiased 2^-17 operand error
    averages out over the reduction, leaving ~1e-6 .. 1e-5 on the outputs (a bf16 convolution: 4e-3)."""
    from gpu_util import rel_err
    r = _split_case(case, parts=2)
    assert rel_err(r["y"], r["y_ref"]) < 2e-5, rel_err(r["y"], r["y_ref"])
    if "dx" in r:
        assert rel_err(r["dx"], r["dx_ref"]) < 2e-5
    assert rel_err(r["dw"], r["dw_ref"]) < 2e-5
    print("split2 errors: y %.2e dx %.2e dw %.2e" % (rel_err(r["y"], r["y_ref"]),
                                                      rel_err(r["dx"], r["dx_ref"]) if "dx" in r else 0,
                                                      rel_err(r["dw"], r["dw_ref"])))


@pytest.mark.parametrize("case", SPLIT_CASES, ids=lambda c: "k%dd%d_%dto%d_%dx%d_B%d" % c)
def test_tc_split_conv_is_fp32_grade(case):
    """fprop / dgrad / wgrad on bf16 [hi|mid|lo] operands against float64: the error must be that of a float32
    convolution (a few 2^-24 times sqrt(K)), three orders of magnitude below a bf16 convolution's 2^-9."""
    from gpu_util import rel_err
    r = _split_case(case)
    tol = 2e-6
    assert rel_err(r["y"], r["y_ref"]) < tol, rel_err(r["y"], r["y_ref"])
    if "dx" in r:
        assert rel_err(r["dx"], r["dx_ref"]) < tol, rel_err(r["dx"], r["dx_ref"])
        assert rel_err(r["dx_acc"], 2 * r["dx_ref"]) < 2 * tol
    assert rel_err(r["dw"], r["dw_ref"]) < 5e-6, rel_err(r["dw"], r["dw_ref"])


@pytest.mark.parametrize("case", [(3, 2, 128, 128, 40, 40, 3), (1, 1, 64, 32, 80, 80, 1), (3, 1, 32, 64, 48, 48, 1)],
                         ids=lambda c: "k%dd%d_%dto%d_%dx%d_B%d" % c)
def test_tc_split_fprop_fused_bn_statistics(case):
    from gpu_util import rel_err
    r = _split_case(case, with_stats=True)
    cout = case[3]
    y = r["y"].astype(np.float64).reshape(-1, cout)
    mean, var = y.mean(0), y.var(0)
    istd = 1 / np.sqrt(var + 1e-5)
    got = r["bnp"]
    assert np.max(np.abs(got[:cout] - mean)) < 1e-5 * max(1.0, np.max(np.abs(mean)))
    assert rel_err(got[cout:2 * cout], istd) < 1e-4
    assert rel_err(got[2 * cout:3 * cout], r["gamma"] * istd) < 1e-4
    assert np.array_equal(got[3 * cout:], r["beta"])


@pytest.mark.parametrize("relu", [1, 0])
@pytest.mark.parametrize("case", [(3, 2, 128, 128, 40, 40, 16), (1, 1, 512, 128, 40, 40, 16), (1, 1, 64, 32, 80, 80, 16),
                                  (3, 1, 32, 32, 80, 80, 3), (1, 1, 1024, 256, 40, 40, 16), (1, 1, 1024, 256, 2, 2, 16),
                                  (3, 1, 64, 64, 40, 40, 16)],
                         ids=lambda c: "k%dd%d_%dto%d_%dx%d_B%d" % c)
def test_tc_fprop_fused_bn_apply(case, relu):
    """conv -> statistics -> grid barrier -> BN apply (+ReLU) from the fp32 TMEM accumulators, one cooperative launch
    (basi_tc_conv_set_bn_apply): the raw output, the published parameters and the activated output must match a
    float64 batch-norm of the float64 convolution within bf16 storage rounding."""
    from basi_b200 import _lib
    from basi_b200._lib import ConvDesc
    from basi_b200.engine import Act
    from gpu_util import bf16_round, call, dev, host, rel_err
    k, d, cin, cout, H, W, B = case
    rng = np.random.RandomState(5)
    x = bf16_round(rng.uniform(-1, 1, (B, H, W, cin)).astype(np.float32) + 0.3)
    w = bf16_round((rng.uniform(-1, 1, (k, k, cin, cout)) / np.sqrt(k * k * cin)).astype(np.float32))
    gamma, beta = rng.uniform(0.5, 1.5, cout).astype(np.float32), rng.uniform(-1, 1, cout).astype(np.float32)
    pad = d * (k - 1) // 2
    desc = ConvDesc(k, k, 1, d, pad, pad, 0)
    bt = torch.bfloat16
    xa = Act(torch.from_numpy(x).to("cuda:0").to(bt).contiguous())
    ya = Act(torch.zeros((B, H, W, cout), dtype=bt, device="cuda:0"))
    oa = Act(torch.full((B, H, W, cout), 1e30, dtype=bt, device="cuda:0"))      # sentinel no BN output can equal
    wd, gd, bd = dev(w), dev(gamma), dev(beta)
    w_io = torch.zeros(k * k * cin * cout, dtype=bt, device="cuda:0")
    w_oi = torch.zeros(k * k * cin * cout, dtype=bt, device="cuda:0")
    call("basi_tc_pack_weights", wd.data_ptr(), w_io.data_ptr(), w_oi.data_ptr(), k * k, cin, cout)
    sums = torch.zeros(2 * cout * 8, dtype=torch.float64, device="cuda:0")
    bnp = torch.zeros(4 * cout, device="cuda:0")
    cnt = torch.zeros(2, dtype=torch.int32, device="cuda:0")
    h = C.c_void_p()
    _lib.call("basi_tc_conv_create", 0, C.byref(desc), xa.ref, ya.ref, w_oi.data_ptr(), None, 0, C.byref(h))
    R = float(B * H * W)
    _lib.call("basi_tc_conv_set_bn_stats", h, sums.data_ptr(), gd.data_ptr(), bd.data_ptr(), C.c_double(R),
              C.c_float(1e-5), bnp.data_ptr(), cnt.data_ptr())
    fused = _lib.load().basi_tc_conv_set_bn_apply(h, oa.ref, relu)
    assert fused == 1, "this shape must take the fused path"
    for rep in range(2):          # twice: the barrier counter keeps counting within a step, sums are re-zeroed
        sums.zero_()
        call("basi_tc_conv_run", h)
    torch.cuda.synchronize()
    y, out, got = host(ya).astype(np.float64), host(oa).astype(np.float64), host(bnp)
    _lib.load().basi_tc_conv_destroy(h)
    xt = torch.from_numpy(x).permute(0, 3, 1, 2).double()
    yref = O.conv2d(xt, torch.from_numpy(w).double(), 1, pad, d).permute(0, 2, 3, 1).numpy()
    assert rel_err(y, yref) < 1e-2
    yr = yref.reshape(-1, cout)
    mean, var = yr.mean(0), yr.var(0)
    istd = 1 / np.sqrt(var + 1e-5)
    assert np.max(np.abs(got[:cout] - mean)) < 5e-3 * max(1.0, np.max(np.abs(mean)))
    assert rel_err(got[cout:2 * cout], istd) < 1e-2
    ref = (yref - mean) * istd * gamma + beta
    if relu:
        ref = np.maximum(ref, 0)
    assert rel_err(out, ref) < 1.5e-2, rel_err(out, ref)
    miss = np.argwhere(out > 1e29)
    assert len(miss) == 0, "unwritten elements: %d, first %s, last %s, images %s, channels %s" % (
        len(miss), miss[0], miss[-1], sorted(set(miss[:, 0]))[:8], sorted(set(miss[:, 3]))[:8])


@pytest.mark.parametrize("relu", [1, 0])
@pytest.mark.parametrize("case", [(3, 2, 128, 128, 40, 40, 16), (1, 1, 128, 512, 40, 40, 16), (1, 1, 32, 128, 80, 80, 16),
                                  (3, 1, 32, 32, 80, 80, 3), (3, 1, 64, 64, 40, 40, 16), (1, 1, 64, 256, 40, 40, 5)],
                         ids=lambda c: "k%dd%d_%dto%d_%dx%d_B%d" % c)
def test_tc_dgrad_fused_bn_backward(case, relu):
    """dgrad + the BN(+ReLU) backward of the layer feeding the convolution in one cooperative launch
    (basi_tc_conv_set_bn_bwd): dx (gradient wrt the raw conv output of that layer), dgamma and dbeta against float64."""
    from basi_b200 import _lib
    from basi_b200._lib import ConvDesc
    from basi_b200.engine import Act
    from gpu_util import bf16_round, call, dev, host, rel_err
    k, d, cin, cout, H, W, B = case
    rng = np.random.RandomState(9)
    xraw = bf16_round(rng.uniform(-1, 1, (B, H, W, cin)).astype(np.float32) * 1.5 + 0.2)     # raw conv output of layer L
    w = bf16_round((rng.uniform(-1, 1, (k, k, cin, cout)) / np.sqrt(k * k * cin)).astype(np.float32))
    dy = bf16_round(rng.uniform(-1, 1, (B, H, W, cout)).astype(np.float32))
    gamma, beta = rng.uniform(0.5, 1.5, cin).astype(np.float32), rng.uniform(-0.5, 0.5, cin).astype(np.float32)
    pad = d * (k - 1) // 2
    desc = ConvDesc(k, k, 1, d, pad, pad, 0)
    bt = torch.bfloat16
    # float64 reference: a = [relu](bn(x)); y = conv(a, w); L = sum(y * dy)
    xt = torch.from_numpy(xraw).permute(0, 3, 1, 2).double().requires_grad_(True)
    gt = torch.from_numpy(gamma).double().requires_grad_(True)
    bt_ = torch.from_numpy(beta).double().requires_grad_(True)
    a = O.batch_norm(xt, gt, bt_, bool(relu))
    yref = O.conv2d(a, torch.from_numpy(w).double(), 1, pad, d)
    (yref * torch.from_numpy(dy).permute(0, 3, 1, 2).double()).sum().backward()
    dx_ref = xt.grad.permute(0, 2, 3, 1).numpy()
    xr = xraw.astype(np.float64).reshape(-1, cin)
    mean, var = xr.mean(0), xr.var(0)
    istd = 1 / np.sqrt(var + 1e-5)
    bnp = np.concatenate([mean, istd, gamma * istd, beta]).astype(np.float32)
    xa = Act(torch.from_numpy(xraw).to("cuda:0").to(bt).contiguous())
    dya = Act(torch.from_numpy(dy).to("cuda:0").to(bt).contiguous())
    dxa = Act(torch.full((B, H, W, cin), 1e30, dtype=bt, device="cuda:0"))
    wd, bnpd = dev(w), dev(bnp)
    w_io = torch.zeros(k * k * cin * cout, dtype=bt, device="cuda:0")
    w_oi = torch.zeros(k * k * cin * cout, dtype=bt, device="cuda:0")
    call("basi_tc_pack_weights", wd.data_ptr(), w_io.data_ptr(), w_oi.data_ptr(), k * k, cin, cout)
    dsums = torch.zeros(2 * cin * 8, dtype=torch.float64, device="cuda:0")
    dgam = torch.zeros(cin, device="cuda:0")
    dbet = torch.zeros(cin, device="cuda:0")
    cnt = torch.zeros(2, dtype=torch.int32, device="cuda:0")
    h = C.c_void_p()
    _lib.call("basi_tc_conv_create", 1, C.byref(desc), dya.ref, dxa.ref, w_io.data_ptr(), None, 0, C.byref(h))
    ok = _lib.load().basi_tc_conv_set_bn_bwd(h, xa.ref, bnpd.data_ptr(), relu, dsums.data_ptr(),
                                             C.c_double(float(B * H * W)), dgam.data_ptr(), dbet.data_ptr(), cnt.data_ptr())
    assert ok == 1, "this shape must take the fused path"
    for rep in range(2):
        dsums.zero_(); dgam.zero_(); dbet.zero_()
        call("basi_tc_conv_run", h)
    torch.cuda.synchronize()
    dx = host(dxa).astype(np.float64)
    _lib.load().basi_tc_conv_destroy(h)
    assert not np.any(np.abs(dx) > 1e29), "unwritten elements"
    assert rel_err(dx, dx_ref) < 2e-2, rel_err(dx, dx_ref)
    assert rel_err(host(dbet), bt_.grad.numpy()) < 1e-2
    assert rel_err(host(dgam), gt.grad.numpy()) < 1e-2


@pytest.mark.gpu
def test_pack_weights_multi_equals_single_layer_packing():
    """basi_tc_pack_weights_multi (persistent grid, layer table in shared memory, 128-bit tile path for channel counts
    that are multiples of 4, scalar path otherwise) writes bit for bit what the single-layer basi_tc_pack_weights
    writes, for several layers of one table incl. ragged tiles, and touches nothing else."""
    from basi_b200._lib import PackEntry
    from gpu_util import call, dev, host
    rng = np.random.RandomState(3)
    layers = [(9, 128, 128), (1, 512, 128), (9, 24, 40), (1, 6, 10), (9, 64, 32), (1, 36, 100)]     # (taps, cin, cout)
    bt = torch.bfloat16
    ents, keep, blocks = [], [], 0
    for taps, cin, cout in layers:
        w = rng.uniform(-1, 1, (taps, cin, cout)).astype(np.float32)
        wd = dev(w)
        n = taps * cin * cout
        bufs = [torch.full((n + 64,), 7.0, dtype=bt, device="cuda:0") for _ in range(4)]    # multi io/oi, single io/oi
        tco, tci = -(-cout // 32), -(-cin // 32)
        ents.append(PackEntry(wd.data_ptr(), bufs[0].data_ptr(), bufs[1].data_ptr(), taps, cin, cout, blocks, tco, tci, 0, 0))
        blocks += taps * tco * tci
        call("basi_tc_pack_weights", wd.data_ptr(), bufs[2].data_ptr(), bufs[3].data_ptr(), taps, cin, cout)
        keep.append((wd, bufs, n))
    arr = (PackEntry * len(ents))(*ents)
    table = torch.from_numpy(np.frombuffer(bytes(arr), dtype=np.uint8).copy()).to("cuda:0")
    call("basi_tc_pack_weights_multi", table.data_ptr(), len(ents), blocks)
    for (taps, cin, cout), (wd, bufs, n) in zip(layers, keep):
        io_m, oi_m, io_s, oi_s = [host(b) for b in bufs]
        assert np.array_equal(io_m[:n], io_s[:n]) and np.array_equal(oi_m[:n], oi_s[:n]), (taps, cin, cout)
        assert np.all(io_m[n:] == 7.0) and np.all(oi_m[n:] == 7.0)                            # guard band intact
        w = host(wd)
        want = torch.from_numpy(w).to(bt).float().numpy()
        assert np.array_equal(io_m[:n].reshape(taps, cin, cout), want)
        assert np.array_equal(oi_m[:n].reshape(taps, cout, cin), want.transpose(0, 2, 1))
