"""Golden vectors for PWC-Net's warp(): executes the reference METHOD itself (models/PWCNet.py:164-204, source
extracted from the file, `.cuda()` dropped so it runs on CPU) on seeded inputs, forward and both gradients.

    python oracle/make_golden_warp.py        (needs /root/reference; writes tests/golden/warp_*.npz)
"""
import os
import re
import warnings

import numpy as np
import torch

REF = os.environ.get("REFERENCE_ROOT", "/root/reference")
GOLD = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")

CASES = {"warp_small": (2, 5, 9, 13, 3.0), "warp_pwc_l6": (1, 12, 6, 20, 1.5), "warp_far": (1, 3, 8, 8, 12.0)}


def reference_warp():
    src = open(os.path.join(REF, "models", "PWCNet.py")).read()
    body = re.search(r"    def warp\(self, x, flo\).*?return output \* mask\n", src, re.S).group(0)
    ns = {"torch": torch, "nn": torch.nn, "Variable": torch.autograd.Variable}
    exec("class _M:\n" + body.replace(".cuda()", ""), ns)
    return ns["_M"]().warp


def main():
    warp = reference_warp()
    warnings.simplefilter("ignore")
    for i, (name, (B, C, H, W, sig)) in enumerate(CASES.items()):
        g = torch.Generator().manual_seed(3000 + i)
        x = torch.randn(B, C, H, W, generator=g).requires_grad_()
        flo = (sig * torch.randn(B, 2, H, W, generator=g)).requires_grad_()
        out = warp(x, flo)
        gout = torch.randn(out.shape, generator=g)
        gx, gf = torch.autograd.grad(out, (x, flo), gout)
        np.savez(os.path.join(GOLD, name + ".npz"), x=x.detach().numpy(), flo=flo.detach().numpy(),
                 out=out.detach().numpy(), gout=gout.numpy(), gx=gx.numpy(), gflo=gf.numpy())
        print(name, tuple(out.shape), float((out == 0).float().mean()))


if __name__ == "__main__":
    main()
