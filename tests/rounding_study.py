"""CPU study (not a test): where does a 16-bit tensor-core path lose its accuracy on this network?

Runs ``oracle.pspnet_forward_rounded`` (float32 arithmetic, bf16 -- or with STUDY_FP16=1 fp16 -- roundings injected
where the CUDA path stores 16-bit values) for several storage policies and prints the per-stage rel-l2 error
against the float64 oracle.  The output is committed in profiles/parity_r02.md; the measured GPU table next to it
comes from tests/parity_table.py.

  python tests/rounding_study.py [S] [F] [B]
"""
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
          "conv5_4_bn", "logits"]


def main():
    S = int(sys.argv[1]) if len(sys.argv) > 1 else 320
    Fn = int(sys.argv[2]) if len(sys.argv) > 2 else 32
    B = int(sys.argv[3]) if len(sys.argv) > 3 else 2
    r16 = torch.float16 if os.environ.get("STUDY_FP16") else torch.bfloat16
    torch.set_num_threads(os.cpu_count() or 1)
    from basi_b200.BAISData import SyntheticData
    sd = SyntheticData(B, (S, S), 8, 21, 1, seed=0)
    img, clicks, lab, cls = sd.next_batch()
    data = np.stack([O.pack_input(img[b], clicks[b]) for b in range(B)])
    params = O.init_params(O.param_specs("1NoClass", 21, 1, Fn), 1, trained_like=True)
    with torch.no_grad():
        ref = O.pspnet_forward_rounded(O.to_torch(params, torch.float64), torch.from_numpy(data).double(), S // 8, ())
        p32 = O.to_torch(params, torch.float32)
        x32 = torch.from_numpy(data).float()
        print("S=%d F=%d B=%d, 16-bit type %s: rel-l2 error vs the float64 oracle" % (S, Fn, B, r16))
        print("%-18s " % "policy" + " ".join("%10s" % s.split("/")[0][-10:] for s in STAGES) + "  mask-agree")
        for name, pol in O.ROUNDING_POLICIES.items():
            out = O.pspnet_forward_rounded(p32, x32, S // 8, pol, r16_dtype=r16)
            errs = [float((out[s].double() - ref[s]).norm() / ref[s].norm()) for s in STAGES]
            agree = float(((out["logits"] > 0) == (ref["logits"] > 0)).float().mean())
            print("%-18s " % name + " ".join("%10.2e" % e for e in errs) + "  %.5f" % agree)


if __name__ == "__main__":
    main()
