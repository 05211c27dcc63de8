"""RAFT with the correlation lookup fused into the motion encoder's first convolution,
WITHOUT editing the reference's files.

`models/raft/raft.py:189` computes `corr = corr_fn(coords1)` and hands it to the update block, whose motion
encoder starts with `cor = F.relu(self.convc1(corr))` (`models/raft/update.py:104,111`).  Inside
`fuse_motion_encoder(net)`:

  * `CorrBlock.__call__` returns a lazy handle (block + coordinates) instead of the (B, 324, H, W) tensor;
  * `encoder.convc1` is wrapped: handed the lazy handle it calls `CorrBlock.lookup_convc1` -- ONE kernel, bias and
    ReLU included (the reference's own `F.relu` on top is then the identity) -- and behaves as the plain
    convolution for anything else.

Everything else in raft.py / update.py runs unmodified.  `return_feat_maps` (raft.py:191-192 clones the lookup)
materialises the tensor through `.clone()`.
"""
import contextlib

import torch.nn as nn

from .. import raft_corr


class LazyLookup:
    """What `corr_fn(coords)` returns inside `fuse_motion_encoder`: the lookup, not yet performed."""

    def __init__(self, block, coords, call):
        self.block, self.coords, self._call = block, coords, call

    def materialize(self):
        return self._call(self.block, self.coords)

    def clone(self):                      # raft.py:191-192 (return_feat_maps)
        return self.materialize().clone()


class FusedConvc1(nn.Module):
    def __init__(self, conv, calls):
        super().__init__()
        self.conv, self.calls = conv, calls

    @property
    def weight(self):
        return self.conv.weight

    @property
    def bias(self):
        return self.conv.bias

    def forward(self, x):
        if isinstance(x, LazyLookup):
            self.calls["fused"] += 1
            return x.block.lookup_convc1(x.coords, self.conv.weight, self.conv.bias, relu=True)
        return self.conv(x)


@contextlib.contextmanager
def fuse_motion_encoder(net):
    """Context manager around `net(image1, image2)` for a reference RAFT instance; yields a call counter."""
    enc = net.update_block.encoder
    calls = {"fused": 0}
    orig_conv, orig_call = enc.convc1, raft_corr.CorrBlock.__call__
    enc.convc1 = FusedConvc1(orig_conv, calls)

    def lazy_call(block, coords):
        if block.compute_spatial:
            return orig_call(block, coords)
        return LazyLookup(block, coords, orig_call)

    raft_corr.CorrBlock.__call__ = lazy_call
    try:
        yield calls
    finally:
        raft_corr.CorrBlock.__call__ = orig_call
        enc.convc1 = orig_conv
