"""Diagnosis helper (not a test): run the forward op by op with a sync after each, then print the error of
every kept stage against the CPU oracle.  usage: python tests/debug_forward.py [fp32|bf16] [simt|tc] [B]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from crog_b200 import _lib as L  # noqa: E402
from crog_b200 import synth  # noqa: E402
from crog_b200.model import CROG  # noqa: E402
from oracle import crog_forward as O  # noqa: E402


def main():
    prec = sys.argv[1] if len(sys.argv) > 1 else "fp32"
    impl = sys.argv[2] if len(sys.argv) > 2 else "tc"
    B = int(sys.argv[3]) if len(sys.argv) > 3 else 1
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    cfg = synth.default_cfg(17)
    sd = synth.make_state_dict(cfg, 0, "perturbed")
    model = CROG(cfg, precision=prec, use_cuda_graph=False)
    model.load_state_dict(sd)
    model = model.cuda()
    if impl == "simt":
        model.gemm_impl = L.IMPL_SIMT
    img, word = synth.make_inputs(B, 17)
    plan = model.plan_for(B, 416, keep=True)
    plan.img.copy_(img.cuda()); plan.word.copy_(word.cuda())
    plan.run_debug()
    print(f"[{prec}/{impl}] all {len(plan.ops)} ops ran, {plan.n_launches} launches, {plan.gemm_flops / B / 1e9:.2f} GFLOP/sample issued")
    maps, inter = O.crog_forward(sd, cfg, img, word, keep=True)
    ref = dict(inter)
    for i, nm in enumerate(("mask", "qua", "sin", "cos", "wid")):
        ref["map_" + nm] = maps[i]
    got = {k: v.interior().cpu() for k, v in plan.keep.items() if k in ref}
    for i, nm in enumerate(("mask", "qua", "sin", "cos", "wid")):
        got["map_" + nm] = plan.out[i].cpu()
    for k in ["stem", "layer1", "layer2", "layer3", "layer4", "c5", "word", "state", "fq_neck", "fq_dec", "map_mask", "map_qua",
              "map_sin", "map_cos", "map_wid"]:
        if k in got:
            w = ref[k]
            if w.dim() == 3:
                w = w.reshape(-1, w.shape[-1])
            g = got[k].reshape(w.shape)
            print(f"  {k:10s} max-abs {float((g - w).abs().max()):.3e}  rel-L2 {float((g - w).norm() / (w.norm() + 1e-12)):.3e}  |ref|max {float(w.abs().max()):.3g}")


if __name__ == "__main__":
    main()
