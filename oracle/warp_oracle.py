"""TEST INFRASTRUCTURE ONLY -- restatement of PWCDCNet.warp (models/PWCNet.py:164-204) with torch ops.

Same operations in the same order as the reference method (mesh grid + flow, normalisation by W-1 / H-1,
`grid_sample` with its default align_corners=False on the map and on an all-ones map, `mask >= 0.0001`, multiply);
only the `.cuda()` calls are replaced by the input's device so that it runs on CPU too.  Pinned against
tests/golden/warp_*.npz, which oracle/make_golden_warp.py produced by executing the reference method itself.
Only tests/ may import this module.
"""
import torch


def warp(x, flo):
    B, _, H, W = x.size()
    xx = torch.arange(0, W).view(1, -1).repeat(H, 1)                      # PWCNet.py:174-178
    yy = torch.arange(0, H).view(-1, 1).repeat(1, W)
    xx = xx.view(1, 1, H, W).repeat(B, 1, 1, 1)
    yy = yy.view(1, 1, H, W).repeat(B, 1, 1, 1)
    grid = torch.cat((xx, yy), 1).float().to(x.device)
    vgrid = grid + flo                                                     # :184
    vgrid = torch.stack((2.0 * vgrid[:, 0] / max(W - 1, 1) - 1.0,          # :189-190 (in-place there)
                         2.0 * vgrid[:, 1] / max(H - 1, 1) - 1.0), 1)
    vgrid = vgrid.permute(0, 2, 3, 1)
    output = torch.nn.functional.grid_sample(x, vgrid, align_corners=False)          # :193 (default)
    mask = torch.nn.functional.grid_sample(torch.ones_like(x), vgrid, align_corners=False)   # :194-195
    mask = (mask >= 0.0001).float()                                        # :204
    return output * mask
