"""GPU script (not a pytest file): per-stage parity table of the CUDA paths against the float64 oracle at the
benchmark shape (S=320, F=32), next to the storage-rounding model of the same policy (oracle.pspnet_forward_rounded).

  python tests/parity_table.py [--batch 4] [--variant 1NoClass|2AddClass] [--precisions bf16,f32] [--out file.md]

Output is committed under profiles/ (parity_r02*.md).
"""
import argparse
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "instance-segment-basi_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)

from oracle import basi_oracle as O  # noqa: E402

STAGES = ["conv1_3_3x3_bn", "conv2_3/relu", "conv3_4/relu", "conv4_8/relu", "conv4_23/relu", "conv5_3/relu",
          "conv5_3_pool1_interp", "conv5_3_pool6_interp", "conv5_4_bn"]


def rel2(a, b):
    a, b = np.asarray(a, np.float64).reshape(-1), np.asarray(b, np.float64).reshape(-1)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300))


def iou(a, b):
    inter, union = np.sum(a & b), np.sum(a | b)
    return 1.0 if union == 0 else float(inter) / float(union)


def oracle_reference(params, data, lab, cls, variant, nseg, P, pw, cw, lr):
    keep = tuple(STAGES)
    return O.train_step(params, data, lab, cls, variant, nseg, P, pw, cw, lr, torch.float64, keep=keep)


def engine_run(variant, nseg, S, F, B, classes, precision, loss, params, data, lab, cls, lr, policy_env=None):
    from basi_b200.BAISPSPNet import PSPNet, Placeholder
    from basi_b200.engine import Engine
    net = PSPNet({'data': Placeholder((None, S, S, 4))}, num_classes=classes, num_segment=nseg, is_training=True,
                 last_pool_size=S // 8, filter_number=F, variant=variant)
    eng = Engine(net, B, precision, True, loss)
    eng.set_params(params)
    eng.feed(data, lab, cls, lr)
    eng.step_device()
    torch.cuda.synchronize()
    out = {s: eng.fetch(s) for s in STAGES}
    out["logits"] = eng.seg_logits.t.cpu().numpy()
    out["loss"] = eng.losses()
    g = eng.get_grads()
    out["grads"] = g
    out["tc_layers"] = eng.tc_layers
    out["storage"] = getattr(eng, "storage_policy", "round1")
    del eng
    torch.cuda.empty_cache()
    return out


def compare(out, ref, variant):
    seg_name = O.VARIANTS[variant][0]
    row = {s: rel2(out[s], ref[s]) for s in STAGES}
    row["logits"] = rel2(out["logits"], ref["seg_logits"])
    a, b = out["logits"], ref["seg_logits"]
    if a.shape[-1] == 1:
        ma, mb = a > 0, b > 0                       # sigmoid > 0.5
    else:
        ma, mb = np.argmax(a, -1) == 1, np.argmax(b, -1) == 1
    row["mask_agree"] = float(np.mean(ma == mb))
    row["mask_iou"] = iou(ma, mb)
    ga = np.concatenate([out["grads"][n].reshape(-1) for n in ref["grads"]]).astype(np.float64)
    gb = np.concatenate([ref["grads"][n].reshape(-1) for n in ref["grads"]]).astype(np.float64)
    row["grad_cos"] = float(ga @ gb / (np.linalg.norm(ga) * np.linalg.norm(gb)))
    row["grad_rel2"] = rel2(ga, gb)
    tot = float(np.linalg.norm(gb))
    worst = sorted(((rel2(out["grads"][n], ref["grads"][n]), float(np.linalg.norm(ref["grads"][n])) / tot, n)
                    for n in ref["grads"] if np.linalg.norm(ref["grads"][n]) > 0), reverse=True)
    row["grad_worst"] = worst[:8]
    # error share: which tensors carry the overall gradient error
    share = sorted(((float(np.linalg.norm(np.asarray(out["grads"][n], np.float64) - ref["grads"][n])) /
                     max(float(np.linalg.norm(ga - gb)), 1e-300), n) for n in ref["grads"]), reverse=True)
    row["grad_err_share"] = share[:8]
    row["loss_rel"] = abs(out["loss"][0] - ref["loss"]) / abs(ref["loss"])
    return row


def model_rows(params, data, P, seg_name, ref_stage, policies, r16=torch.bfloat16):
    rows = {}
    with torch.no_grad():
        p32 = O.to_torch(params, torch.float32)
        x32 = torch.from_numpy(data).float()
        for name in policies:
            out = O.pspnet_forward_rounded(p32, x32, P, O.ROUNDING_POLICIES[name], seg_name, r16_dtype=r16)
            r = {}
            for s in STAGES:
                r[s] = rel2(out[s].permute(0, 2, 3, 1).numpy(), ref_stage[s])
            lg = out["logits"].permute(0, 2, 3, 1).numpy()
            r["logits"] = rel2(lg, ref_stage["seg_logits"])
            if lg.shape[-1] == 1:
                ma, mb = lg > 0, ref_stage["seg_logits"] > 0
            else:
                ma, mb = np.argmax(lg, -1) == 1, np.argmax(ref_stage["seg_logits"], -1) == 1
            r["mask_agree"] = float(np.mean(ma == mb))
            r["mask_iou"] = iou(ma, mb)
            rows[name] = r
    return rows


def fmt_table(title, rows):
    cols = STAGES + ["logits"]
    lines = ["### " + title, "",
             "| path | " + " | ".join(c.replace("_3x3_bn", "").replace("/relu", "").replace("conv5_3_", "").replace("_interp", "") for c in cols) +
             " | mask agree / IoU | grad cos | grad rel-l2 | loss rel |",
             "|---|" + "---|" * (len(cols) + 4)]
    for name, r in rows:
        lines.append("| %s | " % name + " | ".join("%.2e" % r[c] for c in cols) +
                     " | %.4f / %.4f | %s | %s | %s |" % (
                         r["mask_agree"], r["mask_iou"],
                         "%.4f" % r["grad_cos"] if "grad_cos" in r else "-",
                         "%.2e" % r["grad_rel2"] if "grad_rel2" in r else "-",
                         "%.1e" % r["loss_rel"] if "loss_rel" in r else "-"))
    return "\n".join(lines) + "\n"


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=4)
    ap.add_argument("--size", type=int, default=320)
    ap.add_argument("--filters", type=int, default=32)
    ap.add_argument("--variant", default="1NoClass")
    ap.add_argument("--precisions", default="bf16,f32")
    ap.add_argument("--models", default="round1,fused,fused_fp32_trunk,operands_only,weights_only,none")
    ap.add_argument("--out", default=None)
    args = ap.parse_args()
    variant, S, F, B, classes = args.variant, args.size, args.filters, args.batch, 21
    nseg = {"1NoClass": 1, "2AddClass": 1, "4BorderClass": 4, "5COCO": 3}[variant]
    kind = "bce" if nseg == 1 else "softmax"
    pw, cw, lr = 3.0, (0.0 if variant == "1NoClass" else 0.2), 5e-3
    torch.set_num_threads(os.cpu_count() or 1)
    from basi_b200.BAISData import SyntheticData
    sd = SyntheticData(B, (S, S), 8, classes, nseg, seed=0)
    img, clicks, lab, cls = sd.next_batch()
    data = np.stack([O.pack_input(img[b], clicks[b]) for b in range(B)])
    params = O.init_params(O.param_specs(variant, classes, nseg, F), 1, trained_like=True)
    ref = oracle_reference(params, data, lab, cls, variant, nseg, S // 8, pw, cw, lr)
    rows = []
    for prec in args.precisions.split(","):
        out = engine_run(variant, nseg, S, F, B, classes, prec, dict(kind=kind, pos_weight=pw, class_weight=cw),
                         params, data, lab, cls, lr)
        r = compare(out, ref, variant)
        rows.append(("CUDA %s (%d tcgen05 plans, storage %s)" % (prec, out["tc_layers"], out["storage"]), r))
        print("%s: worst gradient tensors (rel-l2, share of |g|, name): %s" % (prec, r["grad_worst"]))
        print("%s: largest shares of the gradient error (share, name): %s" % (prec, r["grad_err_share"]))
    for name, r in model_rows(params, data, S // 8, O.VARIANTS[variant][0], ref, args.models.split(",")).items():
        rows.append(("model (bf16 roundings): %s" % name, r))
    if "f16" in args.precisions.split(","):
        for name, r in model_rows(params, data, S // 8, O.VARIANTS[variant][0], ref,
                                  [m for m in args.models.split(",") if m != "none"], torch.float16).items():
            rows.append(("model (fp16 roundings): %s" % name, r))
    text = fmt_table("%s, S=%d, F=%d, B=%d, trained-like weights: rel-l2 error vs the float64 oracle" % (
        variant, S, F, B), rows)
    print(text)
    if args.out:
        with open(args.out, "a") as f:
            f.write(text + "\n")


if __name__ == "__main__":
    main()
