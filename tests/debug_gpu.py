"""Ad-hoc GPU diagnostics (not a test): locate the wrong elements of d(conv5_3_1x1_increase) on the B=1 case."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import conftest  # noqa
import numpy as np, torch
from oracle import basi_oracle as O
from test_gpu_net import _setup, _engine, _rel

case = ("1NoClass", 1, 320, 8, 1, 21, 3.0, 0.0)
variant, nseg, S, F, B, classes, pw, cw = case
params, img, clicks, data, lab, cls, sigma = _setup(variant, nseg, S, F, B, classes)
eng = _engine(variant, nseg, S, F, B, classes, "f32", dict(kind="bce", pos_weight=pw, class_weight=cw))
eng.set_params(params); eng.feed(data, lab, cls, 5e-3); eng.step_device(); torch.cuda.synchronize()
keep = ("conv5_3_1x1_increase", "conv5_3/relu", "conv5_2/relu")
ref = O.train_step(params, data, lab, cls, variant, nseg, S // 8, pw, cw, 5e-3, torch.float64, keep=keep, keep_grads=True)
A = lambda k: eng._acts[eng.net.layers[k].index]
x = A("conv5_3_1x1_increase"); out = A("conv5_3/relu")
dx = x.grad.t.cpu().numpy()[0]; rdx = ref["d:conv5_3_1x1_increase"][0]
err = np.abs(dx - rdx)
print("dx err max", err.max(), "at", np.unravel_index(err.argmax(), err.shape), "ref |max|", np.abs(rdx).max())
print("per-row(h) max err", np.round(err.max(axis=(1, 2)) / np.abs(rdx).max(), 4))
print("per-col(w) max err", np.round(err.max(axis=(0, 2)) / np.abs(rdx).max(), 4))
ce = err.max(axis=(0, 1)) / np.abs(rdx).max()
print("channels with err>1e-3:", np.where(ce > 1e-3)[0][:40], "count", (ce > 1e-3).sum(), "of", ce.size)
# recompute dx in numpy (float64) from the engine's own tensors
o = out.t.cpu().numpy()[0].astype(np.float64); do = out.grad.t.cpu().numpy()[0].astype(np.float64)
xv = x.t.cpu().numpy()[0].astype(np.float64)
g = params["conv5_3_1x1_increase_bn/conv5_3_1x1_increase_bn/gamma"].astype(np.float64)
dy = do * (o > 0)
mean = xv.mean(axis=(0, 1)); var = xv.var(axis=(0, 1)); istd = 1 / np.sqrt(var + 1e-5)
xh = (xv - mean) * istd
mine = g * istd * (dy - dy.mean(axis=(0, 1)) - xh * (dy * xh).mean(axis=(0, 1)))
print("engine dx vs numpy-from-engine-tensors", _rel(dx, mine), " | numpy-from-engine vs oracle", _rel(mine, rdx))
print("mask agreement engine-out vs oracle-out:", np.mean((o > 0) == (ref["conv5_3/relu"][0] > 0)), "n diff", np.sum((o > 0) != (ref["conv5_3/relu"][0] > 0)))
print("dout agreement", _rel(do, ref["d:conv5_3/relu"][0]))
